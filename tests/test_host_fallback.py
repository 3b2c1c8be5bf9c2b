"""Host-side error handling around the library's speculative (no-pivot) factorisation, without a GPU: the LAPACK-style
``info`` codes (positive: exact zero pivot -> ``LinAlgError``, like ``jnp.linalg.inv`` producing inf/nan would surface in the
reference; negative: a structural shortcut did not apply -> ``AssumptionViolated``) and the repeat with the shortcut off."""
import numpy as np
import pytest
import torch

from jaxhps_b200 import _lib


def test_check_info_maps_codes():
    _lib.check_info(torch.zeros(5, dtype=torch.int32), "ok")  # nothing to report
    with pytest.raises(np.linalg.LinAlgError, match="matrix 2 at column 7"):
        _lib.check_info(torch.tensor([0, 0, 7, 0], dtype=torch.int32), "leaf")
    with pytest.raises(_lib.AssumptionViolated):
        _lib.check_info(torch.tensor([0, -2, 0], dtype=torch.int32), "merge")
    # a singular matrix wins over a violated assumption in the same batch
    with pytest.raises(np.linalg.LinAlgError):
        _lib.check_info(torch.tensor([-1, 3], dtype=torch.int32), "both")


def test_fallback_repeats_once_with_speculation_off(monkeypatch):
    calls = []

    class FakeLib:
        def hps_lu_set_speculative(self, on):
            calls.append(("spec", on))

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())
    state = {"n": 0}

    def stage():
        state["n"] += 1
        calls.append(("run", state["n"]))
        if state["n"] == 1:
            raise _lib.AssumptionViolated("merge level 2")
        return "result"

    assert _lib.with_pivoting_fallback(stage) == "result"
    assert calls == [("run", 1), ("spec", 0), ("run", 2), ("spec", _lib._SPEC_DEFAULT)]


def test_fallback_restores_the_switch_when_the_repeat_fails(monkeypatch):
    calls = []

    class FakeLib:
        def hps_lu_set_speculative(self, on):
            calls.append(on)

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())

    def stage():
        if not calls:
            raise _lib.AssumptionViolated("x")
        raise np.linalg.LinAlgError("singular")

    with pytest.raises(np.linalg.LinAlgError):
        _lib.with_pivoting_fallback(stage)
    assert calls == [0, _lib._SPEC_DEFAULT]


def test_fallback_is_transparent_without_a_violation(monkeypatch):
    monkeypatch.setattr(_lib, "load", lambda: pytest.fail("the library switch must not be touched"))
    assert _lib.with_pivoting_fallback(lambda: 42) == 42


def test_speculation_context_sets_and_restores(monkeypatch):
    calls = []

    class FakeLib:
        def hps_lu_set_speculative(self, on):
            calls.append(on)

    monkeypatch.setattr(_lib, "load", lambda: FakeLib())
    with _lib.speculation(False):
        calls.append("body")
    assert calls == [0, "body", _lib._SPEC_DEFAULT]
