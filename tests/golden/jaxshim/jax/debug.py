def print(*a, **k):  # noqa: A001
    return None
