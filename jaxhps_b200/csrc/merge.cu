// merge_stage (one tree level) and down_pass (one tree level) on the device.
//
// Reference: merge/_uniform_3D_DtN.py:127-541 + merge/_schur_complement.py:293-774 (oct),
// merge/_uniform_2D_DtN.py:206-475 (quad), merge/_schur_complement.py:117-147,182-237 (Schur
// step), down_pass/_uniform_3D_DtN.py:116-246, down_pass/_uniform_2D_DtN.py:125-189.
//
// Differences from the reference's formulation (same results, fewer bytes and flops):
//   * D and -C are gathered straight from the children's T with the exterior unknowns already
//     in the parent's face order, so the reference's final permutation (two more copies of T)
//     disappears and structurally-zero blocks of B are never touched;
//   * S = D^-1 (-C) and g~ = D^-1 (-h_int) come from one pivoted LU with the right-hand sides
//     carried along (lu.cu), in place in the caller's S / g~ buffers — no explicit inverse;
//   * T = A (+) B S only multiplies the 3 (2 in 2D) non-zero m x m blocks per block row of B.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace hps {

namespace {

constexpr int MAXC = 8;   // children per merge
constexpr int MAXF = 6;   // faces per child
constexpr int MAXS = 12;  // interface slots per merge
constexpr int MAXE = 24;  // exterior panels per merge

struct Topo {
  int n_child, n_face, n_slot, n_ext, n_intf;  // n_intf: interior faces per child
  // role[c][f] >= 0: interface slot; < 0: exterior panel -(role+1) in the parent's ordering
  signed char role[MAXC][MAXF];
  unsigned char flip[MAXC][MAXF];       // interface traversed in reverse by this child
  signed char slot_face[MAXC][MAXS];    // face of child c lying on slot s, or -1
  signed char slot_owner[MAXS][2];      // the two children sharing slot s
  signed char ext_child[MAXE], ext_face[MAXE];
  signed char ext_slot[MAXE][3];        // interface slots of the child owning panel e
  // Order in which the exterior panels sit in the right-hand side -C during the solve: sorted by the first
  // interface slot of the owning child (for the octree: the reference's region order a..h, faces ascending), so
  // that the columns with leading zero rows form a suffix at every row (lu.cu, RhsDesc).
  signed char ext_pos[MAXE];            // panel e -> position in the solve order
  signed char ext_at[MAXE];             // position -> panel
  signed char ext_first_slot[MAXE];     // position -> first interface slot of that panel's child
};

// 3D: children a..h, faces x-,x+,y-,y+,z-,z+; interface ids 9..20 -> slots 0..11
// (merge/_uniform_3D_DtN.py:238-380); exterior panels in face order, four panels per face:
// Face0 [e,h,d,a] Face1 [f,g,c,b] Face2 [e,f,b,a] Face3 [h,g,c,d] Face4 [e,f,g,h] Face5 [a,b,c,d]
// (merge/_uniform_3D_DtN.py:507-541).
Topo make_oct() {
  Topo t = {};
  t.n_child = 8; t.n_face = 6; t.n_slot = 12; t.n_ext = 24; t.n_intf = 3;
  const int X = -1;  // exterior
  const int intf[8][6] = {
      {X, 9, X, 12, 17, X},   // a
      {9, X, X, 10, 18, X},   // b
      {11, X, 10, X, 19, X},  // c
      {X, 11, 12, X, 20, X},  // d
      {X, 13, X, 16, X, 17},  // e
      {13, X, X, 14, X, 18},  // f
      {15, X, 14, X, X, 19},  // g
      {X, 15, 16, X, X, 20},  // h
  };
  const int face_children[6][4] = {{4, 7, 3, 0}, {5, 6, 2, 1}, {4, 5, 1, 0}, {7, 6, 2, 3}, {4, 5, 6, 7}, {0, 1, 2, 3}};
  for (int c = 0; c < 8; ++c)
    for (int f = 0; f < 6; ++f) {
      if (intf[c][f] >= 0) {
        t.role[c][f] = (signed char)(intf[c][f] - 9);
      } else {
        int pos = -1;
        for (int k = 0; k < 4; ++k)
          if (face_children[f][k] == c) pos = k;
        t.role[c][f] = (signed char)(-(f * 4 + pos) - 1);
      }
    }
  return t;
}

// 2D: children SW,SE,NE,NW; sides S,E,N,W; interfaces 5:a|b 6:b|c 7:c|d 8:d|a -> slots 0..3; the
// second child listed walks the interface backwards (merge/_uniform_2D_DtN.py:385-437);
// exterior panels after the reference's roll: a.S b.S b.E c.E c.N d.N d.W a.W (:343-346).
Topo make_quad() {
  Topo t = {};
  t.n_child = 4; t.n_face = 4; t.n_slot = 4; t.n_ext = 8; t.n_intf = 2;
  const int X = -1;
  const int intf[4][4] = {{X, 5, 8, X}, {X, X, 6, 5}, {6, X, X, 7}, {8, 7, X, X}};
  const int flipped[4][4] = {{0, 0, 1, 0}, {0, 0, 0, 1}, {1, 0, 0, 0}, {0, 1, 0, 0}};
  const int ext_panel[4][4] = {{0, X, X, 7}, {1, 2, X, X}, {X, 3, 4, X}, {X, X, 5, 6}};
  for (int c = 0; c < 4; ++c)
    for (int f = 0; f < 4; ++f) {
      if (intf[c][f] >= 0) {
        t.role[c][f] = (signed char)(intf[c][f] - 5);
        t.flip[c][f] = (unsigned char)flipped[c][f];
      } else {
        t.role[c][f] = (signed char)(-ext_panel[c][f] - 1);
      }
    }
  return t;
}

void finish(Topo& t) {
  for (int c = 0; c < MAXC; ++c)
    for (int s = 0; s < MAXS; ++s) t.slot_face[c][s] = -1;
  int owners[MAXS] = {0};
  for (int c = 0; c < t.n_child; ++c) {
    int slots[3], ns = 0;
    for (int f = 0; f < t.n_face; ++f)
      if (t.role[c][f] >= 0) {
        const int s = t.role[c][f];
        t.slot_face[c][s] = (signed char)f;
        t.slot_owner[s][owners[s]++] = (signed char)c;
        slots[ns++] = s;
      }
    for (int f = 0; f < t.n_face; ++f)
      if (t.role[c][f] < 0) {
        const int e = -t.role[c][f] - 1;
        t.ext_child[e] = (signed char)c;
        t.ext_face[e] = (signed char)f;
        for (int k = 0; k < 3; ++k) t.ext_slot[e][k] = (signed char)(k < ns ? slots[k] : -1);
      }
  }
  // solve order: stable sort of the panels by (first slot of the child, child, face)
  int key[MAXE], order[MAXE];
  for (int e = 0; e < t.n_ext; ++e) {
    int first = MAXS;
    for (int k = 0; k < t.n_intf; ++k) first = std::min<int>(first, t.ext_slot[e][k]);
    key[e] = (first * MAXC + t.ext_child[e]) * MAXF + t.ext_face[e];
    order[e] = e;
  }
  std::sort(order, order + t.n_ext, [&](int a, int b) { return key[a] < key[b]; });
  for (int pos = 0; pos < t.n_ext; ++pos) {
    const int e = order[pos];
    t.ext_at[pos] = (signed char)e;
    t.ext_pos[e] = (signed char)pos;
    t.ext_first_slot[pos] = (signed char)(key[e] / (MAXC * MAXF));
  }
}

const Topo& oct_topo() { static Topo t = [] { Topo x = make_oct(); finish(x); return x; }(); return t; }
const Topo& quad_topo() { static Topo t = [] { Topo x = make_quad(); finish(x); return x; }(); return t; }

__device__ __forceinline__ int face_index(const Topo& tp, int c, int f, int t, int m) {
  return f * m + (tp.flip[c][f] ? m - 1 - t : t);
}

// ---- gather: D (workspace), S := -C, g~ := -h_int --------------------------------------
// grid: (column tiles, rows = n_slot*m, merges)
// Only exterior columns [ext0, ext0 + n_ext) are produced (S has leading dimension n_ext): the
// whole range for a normal merge, one rank's share for the column-sharded root merge.
__global__ void __launch_bounds__(256) merge_gather_kernel(Topo tp, int m, int n_src, const double* __restrict__ T_in,
                                                           const double* __restrict__ h_in, double* __restrict__ D,
                                                           double* __restrict__ S, double* __restrict__ gt, int ext0,
                                                           int n_ext) {
  const int n_int = tp.n_slot * m, nf = tp.n_face * m;
  const int mg = blockIdx.z;
  for (int row = blockIdx.y; row < n_int; row += gridDim.y) {
  const int s1 = row / m, t1 = row - s1 * m;
  const int64_t child_sz = (int64_t)nf * nf;
  const double* Tm = T_in + (int64_t)mg * tp.n_child * child_sz;
  const int cA = tp.slot_owner[s1][0], cB = tp.slot_owner[s1][1];
  const int fA = tp.slot_face[cA][s1], fB = tp.slot_face[cB][s1];
  const double* rowA = Tm + cA * child_sz + (int64_t)face_index(tp, cA, fA, t1, m) * nf;
  const double* rowB = Tm + cB * child_sz + (int64_t)face_index(tp, cB, fB, t1, m) * nf;
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < n_int + n_ext + n_src; col += gridDim.x * blockDim.x) {
    if (col < n_int) {
      const int s2 = col / m, t2 = col - s2 * m;
      double v = 0.0;
      const int f2A = tp.slot_face[cA][s2], f2B = tp.slot_face[cB][s2];
      if (f2A >= 0) v += rowA[face_index(tp, cA, f2A, t2, m)];
      if (f2B >= 0) v += rowB[face_index(tp, cB, f2B, t2, m)];
      D[((int64_t)mg * n_int + row) * n_int + col] = v;
    } else if (col < n_int + n_ext) {
      const int cl = col - n_int, ce = cl + ext0;
      const int e = ce / m, u = ce - e * m;
      const int c = tp.ext_child[e], f = tp.ext_face[e];
      double v = 0.0;
      if (c == cA) v = rowA[f * m + u];
      else if (c == cB) v = rowB[f * m + u];
      S[((int64_t)mg * n_int + row) * n_ext + cl] = -v;
    } else {
      const int k = col - n_int - n_ext;
      const double* hm = h_in + (int64_t)mg * tp.n_child * nf * n_src;
      const double v = hm[((int64_t)cA * nf + face_index(tp, cA, fA, t1, m)) * n_src + k] +
                       hm[((int64_t)cB * nf + face_index(tp, cB, fB, t1, m)) * n_src + k];
      gt[((int64_t)mg * n_int + row) * n_src + k] = -v;
    }
  }
  }
}

// ---- exterior part: T_out := A scattered (zero elsewhere), h_out := h_ext, and the non-zero
// m x m blocks of B packed as Bg[merge][e][j][m][m] (j-th interface of the child owning e) ----
__device__ __forceinline__ double zero_of(double) { return 0.0; }
__device__ __forceinline__ double2 zero_of(double2) { return make_double2(0.0, 0.0); }

template <typename E>
__global__ void __launch_bounds__(256) merge_ext_kernel(Topo tp, int m, int n_src, const E* __restrict__ T_in,
                                                        const E* __restrict__ h_in, E* __restrict__ T_out,
                                                        E* __restrict__ h_out, E* __restrict__ Bg) {
  const int n_ext = tp.n_ext * m, nf = tp.n_face * m;
  const int mg = blockIdx.z;
  for (int row = blockIdx.y; row < n_ext; row += gridDim.y) {
  const int e1 = row / m, u1 = row - e1 * m;
  const int c = tp.ext_child[e1], f1 = tp.ext_face[e1];
  const int64_t child_sz = (int64_t)nf * nf;
  const E* trow = T_in + ((int64_t)mg * tp.n_child + c) * child_sz + (int64_t)(f1 * m + u1) * nf;
  const int nB = tp.n_intf * m;
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < n_ext + nB + n_src; col += gridDim.x * blockDim.x) {
    if (col < n_ext) {
      const int e2 = col / m, u2 = col - e2 * m;
      const E v = (tp.ext_child[e2] == c) ? trow[tp.ext_face[e2] * m + u2] : zero_of(E());
      T_out[((int64_t)mg * n_ext + row) * n_ext + col] = v;
    } else if (col < n_ext + nB) {
      const int cb = col - n_ext;
      const int j = cb / m, t = cb - j * m;
      const int s = tp.ext_slot[e1][j];
      const int fs = tp.slot_face[c][s];
      Bg[((((int64_t)mg * tp.n_ext + e1) * tp.n_intf + j) * m + u1) * m + t] = trow[face_index(tp, c, fs, t, m)];
    } else {
      const int k = col - n_ext - nB;
      h_out[((int64_t)mg * n_ext + row) * n_src + k] =
          h_in[(((int64_t)mg * tp.n_child + c) * nf + f1 * m + u1) * n_src + k];
    }
  }
  }
}


// ---- tile-per-CTA variants of the two gathers (full exterior range) ------------------------------------
// A CTA owns one m x m block of the output (or a horizontal slice of it): which child, which faces and
// whether the block is structurally zero are decided ONCE per CTA from the topology, so the inner loop is
// a plain row copy — a warp per row, lanes along the row, four rows of loads in flight before the first
// store.  (The element-wise kernels above decode the role of every entry through lane-varying look-ups in
// the kernel-parameter tables, which serialises in the constant cache: 0.22 of the HBM rate in round 1.)
template <typename E>
__device__ __forceinline__ E neg_e(E a);
template <>
__device__ __forceinline__ double neg_e<double>(double a) { return -a; }
template <>
__device__ __forceinline__ double2 neg_e<double2>(double2 a) { return make_double2(-a.x, -a.y); }
template <typename E>
__device__ __forceinline__ E add_e2(E a, E b);
template <>
__device__ __forceinline__ double add_e2<double>(double a, double b) { return a + b; }
template <>
__device__ __forceinline__ double2 add_e2<double2>(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// dst[r][c] = sign * (srcA[r'][c'] (+ srcB[r''][c''])), r in [r_lo, r_hi), c in [0, m); null sources give zeros.
// Row / column reversal (2D interfaces walked backwards by one of the two children) via stride signs.
template <typename E, bool NEGATE>
__device__ __forceinline__ void copy_block(E* __restrict__ dst, int64_t ldd, const E* __restrict__ A, int64_t rsA, int csA,
                                           const E* __restrict__ B, int64_t rsB, int csB, int r_lo, int r_hi, int m) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = r_lo + warp; r < r_hi; r += nw) {
    E* d = dst + (int64_t)r * ldd;
    const E* a = A ? A + (int64_t)r * rsA : nullptr;
    const E* b = B ? B + (int64_t)r * rsB : nullptr;
    for (int c0 = lane; c0 < m; c0 += 128) {
      E v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + 32 * i;
        v[i] = zero_of(E());
        if (c < m) {
          if (a) v[i] = a[(int64_t)c * csA];
          if (b) v[i] = add_e2<E>(v[i], b[(int64_t)c * csB]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + 32 * i;
        if (c < m) d[c] = NEGATE ? neg_e<E>(v[i]) : v[i];
      }
    }
  }
}

// grid: x = output block column (n_slot blocks of D, then n_ext blocks of -C, then one for g~),
//       y = interface slot s1 * row_splits + slice, z = merge
__global__ void __launch_bounds__(256) merge_gather_tile_kernel(Topo tp, int m, int n_src, const double* __restrict__ T_in,
                                                                const double* __restrict__ h_in, double* __restrict__ D,
                                                                double* __restrict__ S, double* __restrict__ gt,
                                                                int row_splits, int sorted) {
  const int n_int = tp.n_slot * m, n_ext = tp.n_ext * m, nf = tp.n_face * m;
  const int mg = blockIdx.z, cb = blockIdx.x;
  const int s1 = blockIdx.y / row_splits, sl = blockIdx.y - s1 * row_splits;
  const int rows_per = (m + row_splits - 1) / row_splits;
  const int r_lo = sl * rows_per, r_hi = min(m, r_lo + rows_per);
  const int64_t child_sz = (int64_t)nf * nf;
  const int cA = tp.slot_owner[s1][0], cB = tp.slot_owner[s1][1];
  const int fA = tp.slot_face[cA][s1], fB = tp.slot_face[cB][s1];
  const double* TA = T_in + ((int64_t)mg * tp.n_child + cA) * child_sz;
  const double* TB = T_in + ((int64_t)mg * tp.n_child + cB) * child_sz;
  // row t1 of this slot inside child X: face row fX*m + t1, walked backwards when the child flips the interface
  const int64_t rsA = tp.flip[cA][fA] ? -(int64_t)nf : (int64_t)nf, rsB = tp.flip[cB][fB] ? -(int64_t)nf : (int64_t)nf;
  const double* rowA0 = TA + (int64_t)(fA * m + (tp.flip[cA][fA] ? m - 1 : 0)) * nf;
  const double* rowB0 = TB + (int64_t)(fB * m + (tp.flip[cB][fB] ? m - 1 : 0)) * nf;
  if (cb < tp.n_slot) {
    const int s2 = cb;
    const int f2A = tp.slot_face[cA][s2], f2B = tp.slot_face[cB][s2];
    const double* A = nullptr; const double* B = nullptr;
    int csA = 1, csB = 1;
    if (f2A >= 0) { const bool fl = tp.flip[cA][f2A]; A = rowA0 + f2A * m + (fl ? m - 1 : 0); csA = fl ? -1 : 1; }
    if (f2B >= 0) { const bool fl = tp.flip[cB][f2B]; B = rowB0 + f2B * m + (fl ? m - 1 : 0); csB = fl ? -1 : 1; }
    double* dst = D + ((int64_t)mg * n_int + s1 * m) * n_int + s2 * m;
    copy_block<double, false>(dst, n_int, A, rsA, csA, B, rsB, csB, r_lo, r_hi, m);
  } else if (cb < tp.n_slot + tp.n_ext) {
    const int e = cb - tp.n_slot;
    const int c = tp.ext_child[e], f = tp.ext_face[e];
    const double* A = nullptr;
    int64_t rs = nf;
    if (c == cA) { A = rowA0 + f * m; rs = rsA; } else if (c == cB) { A = rowB0 + f * m; rs = rsB; }
    double* dst = S + ((int64_t)mg * n_int + s1 * m) * n_ext + (sorted ? tp.ext_pos[e] : e) * m;
    copy_block<double, true>(dst, n_ext, A, rs, 1, nullptr, 0, 1, r_lo, r_hi, m);
  } else {
    const double* hm = h_in + (int64_t)mg * tp.n_child * nf * n_src;
    const int total = (r_hi - r_lo) * n_src;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int t1 = r_lo + idx / n_src, k = idx - (idx / n_src) * n_src;
      const double v = hm[((int64_t)cA * nf + face_index(tp, cA, fA, t1, m)) * n_src + k] +
                       hm[((int64_t)cB * nf + face_index(tp, cB, fB, t1, m)) * n_src + k];
      gt[((int64_t)mg * n_int + s1 * m + t1) * n_src + k] = -v;
    }
  }
}

// grid: x = output block column (n_ext blocks of T_out, n_intf packed blocks of B, one for h_out),
//       y = exterior panel e1 * row_splits + slice, z = merge
template <typename E>
__global__ void __launch_bounds__(256) merge_ext_tile_kernel(Topo tp, int m, int n_src, const E* __restrict__ T_in,
                                                             const E* __restrict__ h_in, E* __restrict__ T_out,
                                                             E* __restrict__ h_out, E* __restrict__ Bg, int row_splits) {
  const int n_ext = tp.n_ext * m, nf = tp.n_face * m;
  const int mg = blockIdx.z, cb = blockIdx.x;
  const int e1 = blockIdx.y / row_splits, sl = blockIdx.y - e1 * row_splits;
  const int rows_per = (m + row_splits - 1) / row_splits;
  const int r_lo = sl * rows_per, r_hi = min(m, r_lo + rows_per);
  const int c = tp.ext_child[e1], f1 = tp.ext_face[e1];
  const int64_t child_sz = (int64_t)nf * nf;
  const E* row0 = T_in + ((int64_t)mg * tp.n_child + c) * child_sz + (int64_t)(f1 * m) * nf;
  if (cb < tp.n_ext) {
    const int e2 = cb;
    const E* A = (tp.ext_child[e2] == c) ? row0 + tp.ext_face[e2] * m : nullptr;
    E* dst = T_out + ((int64_t)mg * n_ext + e1 * m) * n_ext + e2 * m;
    copy_block<E, false>(dst, n_ext, A, nf, 1, nullptr, 0, 1, r_lo, r_hi, m);
  } else if (cb < tp.n_ext + tp.n_intf) {
    const int j = cb - tp.n_ext;
    const int s = tp.ext_slot[e1][j];
    const int fs = tp.slot_face[c][s];
    const bool fl = tp.flip[c][fs];
    const E* A = row0 + fs * m + (fl ? m - 1 : 0);
    E* dst = Bg + (((int64_t)mg * tp.n_ext + e1) * tp.n_intf + j) * m * m;
    copy_block<E, false>(dst, m, A, nf, fl ? -1 : 1, nullptr, 0, 1, r_lo, r_hi, m);
  } else {
    const int total = (r_hi - r_lo) * n_src;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int u1 = r_lo + idx / n_src, k = idx - (idx / n_src) * n_src;
      h_out[((int64_t)mg * n_ext + e1 * m + u1) * n_src + k] =
          h_in[(((int64_t)mg * tp.n_child + c) * nf + f1 * m + u1) * n_src + k];
    }
  }
}

// algorithmic bytes of the two gathers of one merge: every output entry written once, every source entry of
// the children's T read once per use
double gather_bytes(const Topo& tp, int m, int n_src) {
  double blocks = (double)tp.n_slot * (tp.n_slot + tp.n_ext);  // written
  for (int s1 = 0; s1 < tp.n_slot; ++s1) {
    const int cA = tp.slot_owner[s1][0], cB = tp.slot_owner[s1][1];
    for (int s2 = 0; s2 < tp.n_slot; ++s2) blocks += (tp.slot_face[cA][s2] >= 0) + (tp.slot_face[cB][s2] >= 0);
    blocks += 2.0 * (tp.n_face - tp.n_intf);  // exterior faces of the two children sharing the slot
  }
  return 8.0 * (blocks * m * m + 3.0 * tp.n_slot * m * n_src);
}
double ext_bytes(const Topo& tp, int m, int n_src) {
  const double ext_per_child = tp.n_face - tp.n_intf;
  const double blocks = (double)tp.n_ext * tp.n_ext + tp.n_ext * ext_per_child + 2.0 * tp.n_ext * tp.n_intf;
  return 8.0 * (blocks * m * m + 2.0 * tp.n_ext * m * n_src);
}

// ---- down pass scatter: children's boundary vectors from g_ext and g_int --------------------
__global__ void __launch_bounds__(256) down_scatter_kernel(Topo tp, int m, int n_src, const double* __restrict__ g_ext,
                                                           const double* __restrict__ g_int, double* __restrict__ out) {
  const int node = blockIdx.y;
  const int nf = tp.n_face * m;
  const int64_t per_node = (int64_t)tp.n_child * nf * n_src;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < per_node; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % n_src);
    const int64_t r = idx / n_src;
    const int c = (int)(r / nf);
    const int fr = (int)(r - (int64_t)c * nf);
    const int f = fr / m, t = fr - f * m;
    const int role = tp.role[c][f];
    double v;
    if (role < 0) {
      const int e = -role - 1;
      v = g_ext[((int64_t)node * tp.n_ext * m + e * m + t) * n_src + k];
    } else {
      const int tt = tp.flip[c][f] ? m - 1 - t : t;
      v = g_int[((int64_t)node * tp.n_slot * m + role * m + tt) * n_src + k];
    }
    out[(int64_t)node * per_node + idx] = v;
  }
}

// X[b] (rows x cols, leading dim cols) = [I ; 0]: identity in the first `cols` rows
__global__ void set_identity_kernel(double* X, int rows, int cols, int64_t stride) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    X[(int64_t)blockIdx.y * stride + e] = (r == c) ? 1.0 : 0.0;
  }
}
int set_identity(cudaStream_t st, int batch, double* X, int rows, int cols, int64_t stride) {
  const int64_t total = (int64_t)rows * cols;
  set_identity_kernel<<<dim3((unsigned)std::min<int64_t>((total + 255) / 256, 2048), batch), 256, 0, st>>>(X, rows, cols, stride);
  HPS_LAUNCH_CHECK("set_identity_kernel");
  return 0;
}

// Up pass gather (reference up_pass/_uniform_2D_DtN.py:110-173): from the children's outgoing data h
// build h_int (sum of the two children sharing each interface) and h_ext.  ext_shift rotates the
// exterior panels (1 = the reference's pre-roll order).  E = double or double2.
template <typename E>
__device__ __forceinline__ E add_e(E a, E b);
template <>
__device__ __forceinline__ double add_e<double>(double a, double b) { return a + b; }
template <>
__device__ __forceinline__ double2 add_e<double2>(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

template <typename E>
__global__ void up_gather_dtn_kernel(Topo tp, int m, int n_src, const E* __restrict__ h_in, E* __restrict__ h_int,
                                     E* __restrict__ h_ext, int ext_shift) {
  const int node = blockIdx.y;
  const int nf = tp.n_face * m, n_int = tp.n_slot * m, n_ext = tp.n_ext * m;
  const E* hm = h_in + (int64_t)node * tp.n_child * nf * n_src;
  const int64_t total = (int64_t)(n_int + n_ext) * n_src;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % n_src);
    const int r = (int)(e / n_src);
    if (r < n_int) {
      const int s = r / m, t = r - s * m;
      const int cA = tp.slot_owner[s][0], cB = tp.slot_owner[s][1];
      const E a = hm[((int64_t)cA * nf + face_index(tp, cA, tp.slot_face[cA][s], t, m)) * n_src + k];
      const E b = hm[((int64_t)cB * nf + face_index(tp, cB, tp.slot_face[cB][s], t, m)) * n_src + k];
      h_int[((int64_t)node * n_int + r) * n_src + k] = add_e<E>(a, b);
    } else {
      const int re = r - n_int;
      const int ep = re / m, u = re - ep * m;
      const int pos = (ep + ext_shift) % tp.n_ext;
      const int c = tp.ext_child[ep], f = tp.ext_face[ep];
      h_ext[((int64_t)node * n_ext + pos * m + u) * n_src + k] = hm[((int64_t)c * nf + f * m + u) * n_src + k];
    }
  }
}

// ---- S columns from the solve order back to the parent's face order, in place ----------------------
// new panel e = old panel ext_pos[e]: the panel permutation is walked cycle by cycle; a thread owns one column
// offset t of one row in every panel of its cycle, so nothing is shared between threads.
struct PanelCycles {
  int n_cyc;
  signed char start[MAXE + 1];
  signed char elem[MAXE];
};
PanelCycles make_cycles(const Topo& tp) {
  PanelCycles pc = {};
  bool seen[MAXE] = {};
  int at = 0;
  for (int e = 0; e < tp.n_ext; ++e) {
    if (seen[e] || tp.ext_pos[e] == e) continue;
    pc.start[pc.n_cyc++] = (signed char)at;
    for (int d = e; !seen[d]; d = tp.ext_pos[d]) { pc.elem[at++] = (signed char)d; seen[d] = true; }
  }
  pc.start[pc.n_cyc] = (signed char)at;
  return pc;
}
// A row of a panel is cut into chunks of tpr*E elements: tpr threads (a power of two, <= 256) own E elements each, and a CTA
// holds 256/tpr rows, so that small panels (m = 100: tpr 128, E 1) do not idle most of the CTA (ncu r02: the m = 100 launch
// was issue-bound with 100 of 1024 element slots in use).
template <int E>
__global__ void __launch_bounds__(256) panel_cycle_kernel(PanelCycles pc, int m, int chunks, int tpr, double* __restrict__ S,
                                                          int64_t ld, int64_t rows) {
  const int cyc = blockIdx.x / chunks, ch = blockIdx.x - cyc * chunks;
  const int b = pc.start[cyc], e = pc.start[cyc + 1];
  const int rpc = 256 / tpr, r_in = threadIdx.x / tpr, tt = threadIdx.x - r_in * tpr;
  const int t0 = ch * tpr * E + tt;
  for (int64_t row = (int64_t)blockIdx.y * rpc + r_in; row < rows; row += (int64_t)gridDim.y * rpc) {
    double* base = S + row * ld;
    double tmp[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int t = t0 + tpr * i;
      if (t < m) tmp[i] = base[pc.elem[b] * m + t];
    }
    for (int k = b; k + 1 < e; ++k) {
      double v[E];
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int t = t0 + tpr * i;
        if (t < m) v[i] = base[pc.elem[k + 1] * m + t];
      }
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const int t = t0 + tpr * i;
        if (t < m) base[pc.elem[k] * m + t] = v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int t = t0 + tpr * i;
      if (t < m) base[pc.elem[e - 1] * m + t] = tmp[i];
    }
  }
}
int unsort_panels(const Topo& tp, cudaStream_t st, int m, double* S, int64_t ld, int64_t rows) {
  const PanelCycles pc = make_cycles(tp);
  if (pc.n_cyc == 0 || rows <= 0) return 0;
  int moved = 0;  // panels that change place: each is read once and written once
  for (int e = 0; e < tp.n_ext; ++e) moved += tp.ext_pos[e] != e;
  prof_begin(PROF_UNSORT, st, 16.0 * (double)rows * moved * m);
  if (m <= 256) {
    int tpr = 32;
    while (tpr < m) tpr *= 2;
    const int rpc = 256 / tpr;
    dim3 grid(pc.n_cyc, (unsigned)std::min<int64_t>((rows + rpc - 1) / rpc, 65535));
    panel_cycle_kernel<1><<<grid, 256, 0, st>>>(pc, m, 1, tpr, S, ld, rows);
  } else {
    const int tpr = m > 512 ? 256 : 128;
    const int chunks = (m + 4 * tpr - 1) / (4 * tpr), rpc = 256 / tpr;
    dim3 grid(pc.n_cyc * chunks, (unsigned)std::min<int64_t>((rows + rpc - 1) / rpc, 65535));
    panel_cycle_kernel<4><<<grid, 256, 0, st>>>(pc, m, chunks, tpr, S, ld, rows);
  }
  prof_end(PROF_UNSORT, st);
  HPS_LAUNCH_CHECK("panel_cycle_kernel");
  return 0;
}
// HPS_MERGE_STRUCT=0 switches the structured forward substitution (and the panel sort it needs) off
bool merge_structured() {
  static const bool on = [] { const char* e = std::getenv("HPS_MERGE_STRUCT"); return !(e && e[0] == '0'); }();
  return on;
}

size_t merge_ws_bytes(const Topo& tp, int n_merges, int m) {
  const size_t n_int = (size_t)tp.n_slot * m;
  return align_up((size_t)n_merges * n_int * n_int * sizeof(double), 256) +
         align_up((size_t)n_merges * tp.n_ext * tp.n_intf * m * m * sizeof(double), 256) +
         lu_workspace_bytes(n_merges, (int)n_int) + 1024;
}

int merge_level(const Topo& tp, cudaStream_t st, int n_merges, int m, int n_src, const double* T_in,
                const double* h_in, double* S, double* gt, double* T_out, double* h_out, int want_T, void* ws,
                size_t ws_bytes, int* info, double* D_inv = nullptr, double* BD_inv = nullptr) {
  if (n_merges <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  if (n_merges > 65535) return fail_arg(2, "n_merges per call is limited to 65535");
  const int n_int = tp.n_slot * m, n_ext = tp.n_ext * m;
  Arena ar(ws, ws_bytes);
  double* D = ar.take<double>((size_t)n_merges * n_int * n_int);
  double* Bg = ar.take<double>((size_t)n_merges * tp.n_ext * tp.n_intf * m * m);
  if (!D || !Bg) return fail_arg(13, "merge: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;

  const int row_splits = (m + 127) / 128;  // big blocks are cut into slices of <= 128 rows
  const bool sorted = merge_structured() && tp.n_ext <= RHS_MAX_SEG;
  {
    dim3 grid(tp.n_slot + tp.n_ext + 1, tp.n_slot * row_splits, n_merges);
    prof_begin(PROF_GATHER, st, n_merges * gather_bytes(tp, m, n_src));
    merge_gather_tile_kernel<<<grid, 256, 0, st>>>(tp, m, n_src, T_in, h_in, D, S, gt, row_splits, sorted ? 1 : 0);
    prof_end(PROF_GATHER, st);
    HPS_LAUNCH_CHECK("merge_gather_tile_kernel");
  }
  const int64_t sS = (int64_t)n_int * n_ext, sG = (int64_t)n_int * n_src, sDi = (int64_t)n_int * n_int;
  RhsDesc rhs[3] = {{S, n_ext, sS, n_ext}, {gt, n_src, sG, n_src}, {D_inv, n_int, sDi, n_int}};
  if (sorted) {  // -C sits in the solve order: panel at position k is zero above its child's first interface
    rhs[0].n_seg = tp.n_ext;
    rhs[0].seg_cols = m;
    for (int k = 0; k < tp.n_ext; ++k) rhs[0].seg_first_row[k] = tp.ext_first_slot[k] * m;
  }
  if (D_inv) HPS_TRY(set_identity(st, n_merges, D_inv, n_int, n_int, sDi));  // third right-hand side: D^-1 itself
  HPS_TRY(lu_solve(st, n_merges, n_int, D, n_int, sDi, D_inv ? 3 : 2, rhs, lu_ws, lu_ws_bytes, info, LU_NO_PIVOT_EXPECTED));
  if (sorted) HPS_TRY(unsort_panels(tp, st, m, S, n_ext, (int64_t)n_merges * n_int));
  if (!want_T && !BD_inv) return 0;

  {
    dim3 grid(tp.n_ext + tp.n_intf + 1, tp.n_ext * row_splits, n_merges);
    prof_begin(PROF_GATHER, st, n_merges * ext_bytes(tp, m, n_src));
    merge_ext_tile_kernel<double><<<grid, 256, 0, st>>>(tp, m, n_src, T_in, h_in, T_out, h_out, Bg, row_splits);
    prof_end(PROF_GATHER, st);
    HPS_LAUNCH_CHECK("merge_ext_tile_kernel");
  }
  const int64_t sT = (int64_t)n_ext * n_ext, sH = (int64_t)n_ext * n_src;
  const int64_t sB = (int64_t)tp.n_ext * tp.n_intf * m * m;
  for (int e = 0; e < tp.n_ext; ++e)
    for (int j = 0; j < tp.n_intf; ++j) {
      const int s = tp.ext_slot[e][j];
      const double* Ablk = Bg + ((int64_t)e * tp.n_intf + j) * m * m;
      // T_out[panel e rows, :] += B_block * S[slot s rows, :]
      HPS_TRY(dgemm(st, m, n_ext, m, 1.0, Ablk, m, sB, S + (int64_t)s * m * n_ext, n_ext, sS, 1.0,
                    T_out + (int64_t)e * m * n_ext, n_ext, sT, n_merges));
      // h_out[panel e rows] += B_block * g~[slot s rows]
      HPS_TRY(dgemm(st, m, n_src, m, 1.0, Ablk, m, sB, gt + (int64_t)s * m * n_src, n_src, sG, 1.0,
                    h_out + (int64_t)e * m * n_src, n_src, sH, n_merges));
      // (B D^-1)[panel e rows] (+)= B_block * D^-1[slot s rows]   (no-source build)
      if (BD_inv)
        HPS_TRY(dgemm(st, m, n_int, m, 1.0, Ablk, m, sB, D_inv + (int64_t)s * m * n_int, n_int, sDi, j == 0 ? 0.0 : 1.0,
                      BD_inv + (int64_t)e * m * n_int, n_int, (int64_t)n_ext * n_int, n_merges));
    }
  return 0;
}

int down_level(const Topo& tp, cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
               const double* gt, double* g_children, void* ws) {
  if (n_nodes <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  if (n_nodes > 65535) return fail_arg(2, "n_nodes per call is limited to 65535");
  const int n_int = tp.n_slot * m, n_ext = tp.n_ext * m;
  double* g_int = static_cast<double*>(ws);
  // g_int = S g_ext + g~
  HPS_TRY(dgemm_affine(st, n_int, n_src, n_ext, S, n_ext, (int64_t)n_int * n_ext, g_ext, n_src, (int64_t)n_ext * n_src,
                       gt, n_src, (int64_t)n_int * n_src, g_int, n_src, (int64_t)n_int * n_src, n_nodes));
  const int64_t per_node = (int64_t)tp.n_child * tp.n_face * m * n_src;
  dim3 grid((unsigned)std::min<int64_t>((per_node + 255) / 256, 1024), n_nodes);
  down_scatter_kernel<<<grid, 256, 0, st>>>(tp, m, n_src, g_ext, g_int, g_children);
  HPS_LAUNCH_CHECK("down_scatter_kernel");
  return 0;
}


// =====================================================================================
// 2D ItI quad merge (complex128), reference merge/_uniform_2D_ItI.py:182-375 and
// merge/_schur_complement.py:6-41,78-114.  Every interface carries two unknown vectors (the
// incoming impedance data of each adjacent child); the unknowns are ordered as the reference
// RETURNS them, [a5,b5,b6,c6,c7,d7,d8,a8], so no row permutation is needed afterwards.  The
// complex system (I + coupling) X = -[C | h_int] is solved through its real embedding with the
// FP64 LU (the reference forms W = I - D12 D21 and an explicit block inverse instead).
// Complex products use the expanded-operand identity: A X == A_view (M x 2K, interleaved) times
// X2 (2K x 2N real), which lands directly in interleaved complex storage.
// =====================================================================================

struct ItiTopo {
  Topo base;                                    // role >= 0: slot of the child's own unknown on that face
  signed char unk_child[8], unk_face[8];        // unknown u = incoming data of child X on its face
  signed char unk_other[8], unk_other_face[8];  // the neighbour Y across that interface, and Y's face there
};

ItiTopo make_iti() {
  ItiTopo t = {};
  Topo q = make_quad();
  const int intf[4][4] = {{-1, 5, 8, -1}, {-1, -1, 6, 5}, {6, -1, -1, 7}, {8, 7, -1, -1}};
  const int out_child[8] = {0, 1, 1, 2, 2, 3, 3, 0};
  const int out_intf[8] = {5, 5, 6, 6, 7, 7, 8, 8};
  t.base = q;
  t.base.n_slot = 8;
  for (int u = 0; u < 8; ++u) {
    const int X = out_child[u], s = out_intf[u];
    int fX = -1, Y = -1, fY = -1;
    for (int f = 0; f < 4; ++f)
      if (intf[X][f] == s) fX = f;
    for (int c = 0; c < 4; ++c)
      for (int f = 0; f < 4; ++f)
        if (c != X && intf[c][f] == s) { Y = c; fY = f; }
    t.unk_child[u] = (signed char)X; t.unk_face[u] = (signed char)fX;
    t.unk_other[u] = (signed char)Y; t.unk_other_face[u] = (signed char)fY;
    t.base.role[X][fX] = (signed char)u;
  }
  finish(t.base);
  return t;
}
const ItiTopo& iti_topo() { static ItiTopo t = make_iti(); return t; }

// grid: (column tiles, rows = 8m, merges).  Writes the embedded De (16m x 16m) and the stacked
// right-hand sides Cs (16m x 8m), gs (16m x n_src).
__global__ void __launch_bounds__(256) iti_gather_kernel(ItiTopo tp, int m, int n_src, const double2* __restrict__ R_in,
                                                         const double2* __restrict__ h_in, double* __restrict__ De,
                                                         double* __restrict__ Cs, double* __restrict__ gs) {
  const int n = 8 * m, nf = 4 * m;
  const int row = blockIdx.y, mg = blockIdx.z;
  const int u = row / m, t = row - u * m;
  const int Y = tp.unk_other[u], fY = tp.unk_other_face[u];
  const int64_t child_sz = (int64_t)nf * nf;
  const double2* rrow = R_in + ((int64_t)mg * 4 + Y) * child_sz + (int64_t)face_index(tp.base, Y, fY, t, m) * nf;
  double* De_m = De + (int64_t)mg * 4 * n * n;
  double* Cs_m = Cs + (int64_t)mg * 2 * n * n;
  double* gs_m = gs + (int64_t)mg * 2 * n * n_src;
  for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < 2 * n + n_src; col += gridDim.x * blockDim.x) {
    if (col < n) {
      const int u2 = col / m, t2 = col - u2 * m;
      double2 z = make_double2(0.0, 0.0);
      if (tp.unk_child[u2] == Y) z = rrow[face_index(tp.base, Y, tp.unk_face[u2], t2, m)];
      if (col == row) z.x += 1.0;
      De_m[(int64_t)row * 2 * n + col] = z.x;
      De_m[(int64_t)row * 2 * n + n + col] = -z.y;
      De_m[(int64_t)(n + row) * 2 * n + col] = z.y;
      De_m[(int64_t)(n + row) * 2 * n + n + col] = z.x;
    } else if (col < 2 * n) {
      const int ce = col - n;
      const int e = ce / m, uu = ce - e * m;
      double2 z = make_double2(0.0, 0.0);
      if (tp.base.ext_child[e] == Y) z = rrow[tp.base.ext_face[e] * m + uu];
      Cs_m[(int64_t)row * n + ce] = -z.x;
      Cs_m[(int64_t)(n + row) * n + ce] = -z.y;
    } else {
      const int k = col - 2 * n;
      const double2 z = h_in[(((int64_t)mg * 4 + Y) * nf + face_index(tp.base, Y, fY, t, m)) * n_src + k];
      gs_m[(int64_t)row * n_src + k] = -z.x;
      gs_m[(int64_t)(n + row) * n_src + k] = -z.y;
    }
  }
}

struct IntPos { int pos[8]; };

// Up pass gather for ItI (reference up_pass/_uniform_2D_ItI.py:139-219): the equation of unknown u is
// driven by the OTHER child's outgoing data; block u is written at position pos[u] of h_int.
__global__ void up_gather_iti_kernel(ItiTopo tp, int m, int n_src, const double2* __restrict__ h_in,
                                     double2* __restrict__ h_int, double2* __restrict__ h_ext, int ext_shift, IntPos ip) {
  const int node = blockIdx.y;
  const int nf = 4 * m, n = 8 * m;
  const double2* hm = h_in + (int64_t)node * 4 * nf * n_src;
  const int64_t total = (int64_t)2 * n * n_src;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % n_src);
    const int r = (int)(e / n_src);
    if (r < n) {
      const int u = r / m, t = r - u * m;
      const int Y = tp.unk_other[u], fY = tp.unk_other_face[u];
      h_int[((int64_t)node * n + ip.pos[u] * m + t) * n_src + k] =
          hm[((int64_t)Y * nf + face_index(tp.base, Y, fY, t, m)) * n_src + k];
    } else {
      const int re = r - n;
      const int ep = re / m, uu = re - ep * m;
      const int pos = (ep + ext_shift) % 8;
      const int c = tp.base.ext_child[ep], f = tp.base.ext_face[ep];
      h_ext[((int64_t)node * n + pos * m + uu) * n_src + k] = hm[((int64_t)c * nf + f * m + uu) * n_src + k];
    }
  }
}

size_t merge_iti_ws_bytes_impl(int n_merges, int m, int n_src) {
  const size_t n = 8 * (size_t)m;
  return align_up((size_t)n_merges * 4 * n * n * 8, 256) + align_up((size_t)n_merges * 2 * n * n * 8, 256) +
         align_up((size_t)n_merges * 2 * n * n_src * 8, 256) + align_up((size_t)n_merges * 4 * n * n * 8, 256) +
         align_up((size_t)n_merges * 4 * n * n_src * 8, 256) + align_up((size_t)n_merges * 8 * 2 * m * m * 16, 256) +
         align_up((size_t)n_merges * 2 * n * n * 8, 256) + align_up((size_t)n_merges * 4 * n * n * 8, 256) +
         lu_workspace_bytes(n_merges, (int)(2 * n)) + 1024;
}


// ---- multi-GPU root merge sharded by child ------------------------------------------------------
// pack: for each local subtree root c (child index child0 + c of the root merge) extract
//   Dblk[c] = T_c[int faces, int faces] (3m x 3m), Cblk[c] = T_c[int faces, ext faces] (3m x 3m),
//   hblk[c] = h_c[int faces]; interior faces in ascending interface-slot order of that child,
//   exterior faces in ascending face order.
__global__ void __launch_bounds__(256) root_pack_kernel(Topo tp, int m, int n_src, int child0, const double* __restrict__ T,
                                                        const double* __restrict__ h, double* __restrict__ Dblk,
                                                        double* __restrict__ Cblk, double* __restrict__ hblk) {
  const int c_loc = blockIdx.z, c = child0 + c_loc;
  const int nf = 6 * m, n3 = 3 * m;
  int int_face[3], ext_face[3], ni = 0, ne = 0;
  // slots ascending: ext_slot of any exterior panel of this child lists them in face order; sort by slot
  for (int f = 0; f < 6; ++f) {
    if (tp.role[c][f] >= 0) int_face[ni++] = f; else ext_face[ne++] = f;
  }
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (tp.role[c][int_face[b]] < tp.role[c][int_face[a]]) { const int t = int_face[a]; int_face[a] = int_face[b]; int_face[b] = t; }
  const double* Tc = T + (int64_t)c_loc * nf * nf;
  for (int row = blockIdx.y; row < n3; row += gridDim.y) {
    const int i = row / m, t1 = row - i * m;
    const double* trow = Tc + (int64_t)(int_face[i] * m + t1) * nf;
    for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < 2 * n3 + n_src; col += gridDim.x * blockDim.x) {
      if (col < n3) {
        const int j = col / m, t2 = col - j * m;
        Dblk[((int64_t)c_loc * n3 + row) * n3 + col] = trow[int_face[j] * m + t2];
      } else if (col < 2 * n3) {
        const int cc = col - n3;
        const int j = cc / m, t2 = cc - j * m;
        Cblk[((int64_t)c_loc * n3 + row) * n3 + cc] = trow[ext_face[j] * m + t2];
      } else {
        const int k = col - 2 * n3;
        hblk[((int64_t)c_loc * n3 + row) * n_src + k] = h[((int64_t)c_loc * nf + int_face[i] * m + t1) * n_src + k];
      }
    }
  }
}

// assemble: D (12m x 12m) from all 8 children's Dblk, C_r = -[C columns of the local children]
// (12m x 3m*n_local), g~ := -h_int.
__global__ void __launch_bounds__(256) root_assemble_kernel(Topo tp, int m, int n_src, int child0, int n_local,
                                                            const double* __restrict__ Dblk_all,
                                                            const double* __restrict__ hblk_all,
                                                            const double* __restrict__ Cblk_loc, double* __restrict__ D,
                                                            double* __restrict__ Cr, double* __restrict__ gt) {
  const int n_int = 12 * m, n3 = 3 * m, ncr = n3 * n_local;
  // local index (0..2) of slot s inside child c = number of that child's slots below s
  auto loc = [&](int c, int s) {
    int k = 0;
    for (int s2 = 0; s2 < s; ++s2) k += (tp.slot_face[c][s2] >= 0);
    return k;
  };
  for (int row = blockIdx.y; row < n_int; row += gridDim.y) {
    const int s1 = row / m, t1 = row - s1 * m;
    const int cA = tp.slot_owner[s1][0], cB = tp.slot_owner[s1][1];
    const int iA = loc(cA, s1), iB = loc(cB, s1);
    const double* rowA = Dblk_all + ((int64_t)cA * n3 + iA * m + t1) * n3;
    const double* rowB = Dblk_all + ((int64_t)cB * n3 + iB * m + t1) * n3;
    for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < n_int + ncr + n_src; col += gridDim.x * blockDim.x) {
      if (col < n_int) {
        const int s2 = col / m, t2 = col - s2 * m;
        double v = 0.0;
        if (tp.slot_face[cA][s2] >= 0) v += rowA[loc(cA, s2) * m + t2];
        if (tp.slot_face[cB][s2] >= 0) v += rowB[loc(cB, s2) * m + t2];
        D[(int64_t)row * n_int + col] = v;
      } else if (col < n_int + ncr) {
        const int cc = col - n_int;
        const int cl = cc / n3, rest = cc - cl * n3;
        const int c = child0 + cl;
        double v = 0.0;
        if (c == cA) v = Cblk_loc[((int64_t)cl * n3 + iA * m + t1) * n3 + rest];
        else if (c == cB) v = Cblk_loc[((int64_t)cl * n3 + iB * m + t1) * n3 + rest];
        Cr[(int64_t)row * ncr + cc] = -v;
      } else {
        const int k = col - n_int - ncr;
        gt[(int64_t)row * n_src + k] = -(hblk_all[((int64_t)cA * n3 + iA * m + t1) * n_src + k] +
                                         hblk_all[((int64_t)cB * n3 + iB * m + t1) * n_src + k]);
      }
    }
  }
}

// The same assembly for an arbitrary set of exterior PANELS (one panel = one exterior face of one child, m columns):
// the balanced column distribution of the multi-GPU root.  Cpan[k] (3m x m) = T_c[int faces (slot order), that face]
// of the child c = pan.child[k]; columns of C_r follow the order of the panels.
struct PanelList { int n; signed char child[MAXE]; };
__global__ void __launch_bounds__(256) root_assemble_panels_kernel(Topo tp, int m, int n_src, PanelList pan,
                                                                   const double* __restrict__ Dblk_all,
                                                                   const double* __restrict__ hblk_all,
                                                                   const double* __restrict__ Cpan, double* __restrict__ D,
                                                                   double* __restrict__ Cr, double* __restrict__ gt) {
  const int n_int = 12 * m, n3 = 3 * m, ncr = pan.n * m;
  auto loc = [&](int c, int s) {
    int k = 0;
    for (int s2 = 0; s2 < s; ++s2) k += (tp.slot_face[c][s2] >= 0);
    return k;
  };
  for (int row = blockIdx.y; row < n_int; row += gridDim.y) {
    const int s1 = row / m, t1 = row - s1 * m;
    const int cA = tp.slot_owner[s1][0], cB = tp.slot_owner[s1][1];
    const int iA = loc(cA, s1), iB = loc(cB, s1);
    const double* rowA = Dblk_all + ((int64_t)cA * n3 + iA * m + t1) * n3;
    const double* rowB = Dblk_all + ((int64_t)cB * n3 + iB * m + t1) * n3;
    for (int col = blockIdx.x * blockDim.x + threadIdx.x; col < n_int + ncr + n_src; col += gridDim.x * blockDim.x) {
      if (col < n_int) {
        const int s2 = col / m, t2 = col - s2 * m;
        double v = 0.0;
        if (tp.slot_face[cA][s2] >= 0) v += rowA[loc(cA, s2) * m + t2];
        if (tp.slot_face[cB][s2] >= 0) v += rowB[loc(cB, s2) * m + t2];
        D[(int64_t)row * n_int + col] = v;
      } else if (col < n_int + ncr) {
        const int cc = col - n_int;
        const int k = cc / m, t2 = cc - k * m;
        const int c = pan.child[k];
        double v = 0.0;
        if (c == cA) v = Cpan[((int64_t)k * n3 + iA * m + t1) * m + t2];
        else if (c == cB) v = Cpan[((int64_t)k * n3 + iB * m + t1) * m + t2];
        Cr[(int64_t)row * ncr + cc] = -v;
      } else {
        const int k = col - n_int - ncr;
        gt[(int64_t)row * n_src + k] = -(hblk_all[((int64_t)cA * n3 + iA * m + t1) * n_src + k] +
                                         hblk_all[((int64_t)cB * n3 + iB * m + t1) * n_src + k]);
      }
    }
  }
}
int first_slot_of_child(const Topo& tp, int c) {
  int first = 0;
  while (tp.slot_face[c][first] < 0) ++first;
  return first;
}
int make_panel_list(int n_panels, const int* panel_child, PanelList& pan) {
  if (n_panels <= 0 || n_panels > MAXE || !panel_child) return fail_arg(4, "1..24 panels expected");
  pan.n = n_panels;
  for (int k = 0; k < n_panels; ++k) {
    if (panel_child[k] < 0 || panel_child[k] > 7) return fail_arg(5, "panel child out of range");
    pan.child[k] = (signed char)panel_child[k];
  }
  return 0;
}

// Column-sharded merge for the multi-GPU root: S[:, ext0:ext0+ncols] and g~ from the children's T.
int merge_cols(const Topo& tp, cudaStream_t st, int m, int n_src, const double* T_in, const double* h_in, int ext0,
               int ncols, double* S_cols, double* gt, void* ws, size_t ws_bytes, int* info) {
  if (m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  const int n_int = tp.n_slot * m, n_ext = tp.n_ext * m;
  if (ext0 < 0 || ncols <= 0 || ext0 + ncols > n_ext) return fail_arg(6, "column window out of range");
  Arena ar(ws, ws_bytes);
  double* D = ar.take<double>((size_t)n_int * n_int);
  if (!D) return fail_arg(11, "merge_cols: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  const int cols = n_int + ncols + n_src;
  dim3 grid(std::min((cols + 255) / 256, 64), std::min(n_int, 65535), 1);
  merge_gather_kernel<<<grid, 256, 0, st>>>(tp, m, n_src, T_in, h_in, D, S_cols, gt, ext0, ncols);
  HPS_LAUNCH_CHECK("merge_gather_kernel");
  RhsDesc rhs[2] = {{S_cols, ncols, (int64_t)n_int * ncols, ncols}, {gt, n_src, (int64_t)n_int * n_src, n_src}};
  return lu_solve(st, 1, n_int, D, n_int, (int64_t)n_int * n_int, 2, rhs, lu_ws, lu_ws_bytes, info, LU_NO_PIVOT_EXPECTED);
}

int down_scatter(const Topo& tp, cudaStream_t st, int n_nodes, int m, int n_src, const double* g_ext,
                 const double* g_int, double* g_children) {
  if (n_nodes <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  const int64_t per_node = (int64_t)tp.n_child * tp.n_face * m * n_src;
  dim3 grid((unsigned)std::min<int64_t>((per_node + 255) / 256, 1024), n_nodes);
  down_scatter_kernel<<<grid, 256, 0, st>>>(tp, m, n_src, g_ext, g_int, g_children);
  HPS_LAUNCH_CHECK("down_scatter_kernel");
  return 0;
}

}  // namespace


int root_pack_oct(cudaStream_t st, int n_local, int child0, int m, int n_src, const double* T, const double* h,
                  double* Dblk, double* Cblk, double* hblk) {
  if (n_local <= 0 || m <= 0 || n_src <= 0 || child0 < 0 || child0 + n_local > 8) return fail_arg(2, "bad child range");
  const int cols = 6 * m + n_src;
  dim3 grid(std::min((cols + 255) / 256, 64), std::min(3 * m, 65535), n_local);
  root_pack_kernel<<<grid, 256, 0, st>>>(oct_topo(), m, n_src, child0, T, h, Dblk, Cblk, hblk);
  HPS_LAUNCH_CHECK("root_pack_kernel");
  return 0;
}

// assembly only (distributed factorisation path): D (12m x 12m), S_r := -C_r, g~ := -h_int
int root_assemble_oct(cudaStream_t st, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                      const double* hblk_all, const double* Cblk_loc, double* D, double* S_r, double* gt) {
  if (m <= 0 || n_src <= 0 || n_local <= 0 || child0 < 0 || child0 + n_local > 8) return fail_arg(2, "bad child range");
  const int n_int = 12 * m, ncr = 3 * m * n_local;
  const int cols = n_int + ncr + n_src;
  dim3 grid(std::min((cols + 255) / 256, 64), std::min(n_int, 65535), 1);
  root_assemble_kernel<<<grid, 256, 0, st>>>(oct_topo(), m, n_src, child0, n_local, Dblk_all, hblk_all, Cblk_loc, D, S_r, gt);
  HPS_LAUNCH_CHECK("root_assemble_kernel");
  return 0;
}

size_t root_solve_oct_ws_bytes(int m) {
  const size_t n_int = 12 * (size_t)m;
  return align_up(n_int * n_int * sizeof(double), 256) + lu_workspace_bytes(1, (int)n_int) + 1024;
}

// S_r (12m x 3m*n_local) = columns of the root S belonging to the local children's exterior faces
// (child-major, faces ascending), g~ (12m x n_src).
int root_solve_oct(cudaStream_t st, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                   const double* hblk_all, const double* Cblk_loc, double* S_r, double* gt, void* ws, size_t ws_bytes,
                   int* info) {
  if (m <= 0 || n_src <= 0 || n_local <= 0 || child0 < 0 || child0 + n_local > 8) return fail_arg(2, "bad child range");
  const int n_int = 12 * m, ncr = 3 * m * n_local;
  Arena ar(ws, ws_bytes);
  double* D = ar.take<double>((size_t)n_int * n_int);
  if (!D) return fail_arg(11, "root_solve: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  const int cols = n_int + ncr + n_src;
  dim3 grid(std::min((cols + 255) / 256, 64), std::min(n_int, 65535), 1);
  root_assemble_kernel<<<grid, 256, 0, st>>>(oct_topo(), m, n_src, child0, n_local, Dblk_all, hblk_all, Cblk_loc, D, S_r, gt);
  HPS_LAUNCH_CHECK("root_assemble_kernel");
  RhsDesc rhs[2] = {{S_r, ncr, (int64_t)n_int * ncr, ncr}, {gt, n_src, (int64_t)n_int * n_src, n_src}};
  if (merge_structured()) root_cols_structure(child0, n_local, m, rhs[0].n_seg, rhs[0].seg_cols, rhs[0].seg_first_row);
  return lu_solve(st, 1, n_int, D, n_int, (int64_t)n_int * n_int, 2, rhs, lu_ws, lu_ws_bytes, info, LU_NO_PIVOT_EXPECTED);
}

// Leading-zero structure of the root's -C_r (columns of children child0 .. child0+n_local-1, child-major: already
// sorted by first interface): 3 segments of m columns per child, zero above the child's first interface.
void root_cols_structure(int child0, int n_local, int m, int& n_seg, int& seg_cols, int* seg_first_row) {
  const Topo& tp = oct_topo();
  n_seg = 3 * n_local;
  seg_cols = m;
  for (int cl = 0; cl < n_local; ++cl) {
    int first = 0;
    while (tp.slot_face[child0 + cl][first] < 0) ++first;
    for (int k = 0; k < 3; ++k) seg_first_row[3 * cl + k] = first * m;
  }
}

// Panel versions (balanced multi-GPU root): panel_child[k] = child owning the k-th panel of this rank.
int root_assemble_panels(cudaStream_t st, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                         const double* hblk_all, const double* Cpan, double* D, double* S_r, double* gt) {
  if (m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  PanelList pan;
  HPS_TRY(make_panel_list(n_panels, panel_child, pan));
  const int n_int = 12 * m, cols = n_int + n_panels * m + n_src;
  dim3 grid(std::min((cols + 255) / 256, 64), std::min(n_int, 65535), 1);
  root_assemble_panels_kernel<<<grid, 256, 0, st>>>(oct_topo(), m, n_src, pan, Dblk_all, hblk_all, Cpan, D, S_r, gt);
  HPS_LAUNCH_CHECK("root_assemble_panels_kernel");
  return 0;
}
// leading-zero structure of such a panel set; fails unless the panels are sorted by their child's first interface
int root_panels_structure(int n_panels, const int* panel_child, int m, int& n_seg, int& seg_cols, int* seg_first_row) {
  PanelList pan;
  HPS_TRY(make_panel_list(n_panels, panel_child, pan));
  const Topo& tp = oct_topo();
  n_seg = n_panels;
  seg_cols = m;
  for (int k = 0; k < n_panels; ++k) {
    seg_first_row[k] = first_slot_of_child(tp, pan.child[k]) * m;
    if (k > 0 && seg_first_row[k] < seg_first_row[k - 1]) return fail_arg(2, "panels must be sorted by first interface");
  }
  return 0;
}
int root_solve_panels(cudaStream_t st, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                      const double* hblk_all, const double* Cpan, double* S_r, double* gt, void* ws, size_t ws_bytes,
                      int* info) {
  if (m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  const int n_int = 12 * m, ncr = n_panels * m;
  Arena ar(ws, ws_bytes);
  double* D = ar.take<double>((size_t)n_int * n_int);
  if (!D) return fail_arg(11, "root_solve: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  HPS_TRY(root_assemble_panels(st, m, n_src, n_panels, panel_child, Dblk_all, hblk_all, Cpan, D, S_r, gt));
  RhsDesc rhs[2] = {{S_r, ncr, (int64_t)n_int * ncr, ncr}, {gt, n_src, (int64_t)n_int * n_src, n_src}};
  if (merge_structured()) {
    int first[RHS_MAX_SEG];
    if (root_panels_structure(n_panels, panel_child, m, rhs[0].n_seg, rhs[0].seg_cols, first) == 0)
      for (int k = 0; k < n_panels; ++k) rhs[0].seg_first_row[k] = first[k];
    else
      rhs[0].n_seg = 0;  // unsorted panels: solve without the shortcut
  }
  return lu_solve(st, 1, n_int, D, n_int, (int64_t)n_int * n_int, 2, rhs, lu_ws, lu_ws_bytes, info, LU_NO_PIVOT_EXPECTED);
}

int merge_oct_root_cols(cudaStream_t st, int m, int n_src, const double* T_in, const double* h_in, int ext0, int ncols,
                        double* S_cols, double* gt, void* ws, size_t ws_bytes, int* info) {
  return merge_cols(oct_topo(), st, m, n_src, T_in, h_in, ext0, ncols, S_cols, gt, ws, ws_bytes, info);
}
int down_oct_scatter(cudaStream_t st, int n_nodes, int m, int n_src, const double* g_ext, const double* g_int,
                     double* g_children) {
  return down_scatter(oct_topo(), st, n_nodes, m, n_src, g_ext, g_int, g_children);
}


// helpers defined in leaf.cu
int stacked_to_complex(cudaStream_t st, int batch, int rows, int cols, const double* Xs, int64_t sXs, double* X,
                       int64_t sX, double* X2, int64_t sX2);
int complex_expand(cudaStream_t st, int batch, int rows, int cols, const double* X, int64_t sX, double* X2, int64_t sX2,
                   double alpha = 1.0);

size_t merge_quad_iti_ws_bytes(int n_merges, int m, int n_src) { return merge_iti_ws_bytes_impl(n_merges, m, n_src); }

// R_in [4n][4m][4m], h_in [4n][4m][n_src]; S [n][8m][8m], gt [n][8m][n_src], R_out [n][8m][8m], h_out [n][8m][n_src];
// all complex128 stored interleaved.
int merge_quad_iti_level(cudaStream_t st, int n_merges, int m, int n_src, const double* R_in, const double* h_in,
                         double* S, double* gt, double* R_out, double* h_out, int want_T, void* ws, size_t ws_bytes,
                         int* info, double* D_inv, double* BD_inv) {
  if (n_merges <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  if (n_merges > 65535) return fail_arg(2, "n_merges per call is limited to 65535");
  const ItiTopo& tp = iti_topo();
  const int n = 8 * m, n2 = 2 * n;
  Arena ar(ws, ws_bytes);
  double* De = ar.take<double>((size_t)n_merges * n2 * n2);
  double* Cs = ar.take<double>((size_t)n_merges * n2 * n);
  double* gs = ar.take<double>((size_t)n_merges * n2 * n_src);
  double* S2 = ar.take<double>((size_t)n_merges * n2 * n2);
  double* g2 = ar.take<double>((size_t)n_merges * n2 * 2 * n_src);
  double* Bg = ar.take<double>((size_t)n_merges * 8 * 2 * m * m * 2);
  double* Dis = ar.take<double>((size_t)n_merges * n2 * n);   // stacked D^-1 (no-source build)
  double* Di2 = ar.take<double>((size_t)n_merges * n2 * n2);  // its expanded form
  if (!De || !Cs || !gs || !S2 || !g2 || !Bg || !Dis || !Di2) return fail_arg(13, "merge_iti: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  {
    const int cols = 2 * n + n_src;
    dim3 grid(std::min((cols + 255) / 256, 64), n, n_merges);
    iti_gather_kernel<<<grid, 256, 0, st>>>(tp, m, n_src, reinterpret_cast<const double2*>(R_in),
                                            reinterpret_cast<const double2*>(h_in), De, Cs, gs);
    HPS_LAUNCH_CHECK("iti_gather_kernel");
  }
  RhsDesc rhs[3] = {{Cs, n, (int64_t)n2 * n, n}, {gs, n_src, (int64_t)n2 * n_src, n_src}, {Dis, n, (int64_t)n2 * n, n}};
  if (D_inv) HPS_TRY(set_identity(st, n_merges, Dis, n2, n, (int64_t)n2 * n));  // stacked [I ; 0]
  HPS_TRY(lu_solve(st, n_merges, n2, De, n2, (int64_t)n2 * n2, D_inv ? 3 : 2, rhs, lu_ws, lu_ws_bytes, info));
  if (D_inv) HPS_TRY(stacked_to_complex(st, n_merges, n, n, Dis, (int64_t)n2 * n, D_inv, (int64_t)n * n, Di2, (int64_t)n2 * n2));
  HPS_TRY(stacked_to_complex(st, n_merges, n, n, Cs, (int64_t)n2 * n, S, (int64_t)n * n, S2, (int64_t)n2 * n2));
  HPS_TRY(stacked_to_complex(st, n_merges, n, n_src, gs, (int64_t)n2 * n_src, gt, (int64_t)n * n_src, g2,
                             (int64_t)n2 * 2 * n_src));
  if (!want_T && !BD_inv) return 0;
  {
    const int row_splits = (m + 127) / 128;
    dim3 grid(tp.base.n_ext + tp.base.n_intf + 1, tp.base.n_ext * row_splits, n_merges);
    merge_ext_tile_kernel<double2><<<grid, 256, 0, st>>>(
        tp.base, m, n_src, reinterpret_cast<const double2*>(R_in), reinterpret_cast<const double2*>(h_in),
        reinterpret_cast<double2*>(R_out), reinterpret_cast<double2*>(h_out), reinterpret_cast<double2*>(Bg), row_splits);
    HPS_LAUNCH_CHECK("merge_ext_tile_kernel<complex>");
  }
  // R_out[panel e] += B_block (complex m x m) * S[unknown slot rows]; real GEMMs on the interleaved views
  const int64_t sB = (int64_t)8 * 2 * m * m * 2, sS2 = (int64_t)n2 * n2, sR = (int64_t)n * n * 2;
  const int64_t sg2 = (int64_t)n2 * 2 * n_src, sH = (int64_t)n * n_src * 2;
  for (int e = 0; e < 8; ++e)
    for (int j = 0; j < 2; ++j) {
      const int slot = tp.base.ext_slot[e][j];
      const double* Ablk = Bg + ((int64_t)e * 2 + j) * m * m * 2;
      HPS_TRY(dgemm(st, m, 2 * n, 2 * m, 1.0, Ablk, 2 * m, sB, S2 + (int64_t)2 * slot * m * n2, n2, sS2, 1.0,
                    R_out + (int64_t)e * m * n2, n2, sR, n_merges));
      HPS_TRY(dgemm(st, m, 2 * n_src, 2 * m, 1.0, Ablk, 2 * m, sB, g2 + (int64_t)2 * slot * m * 2 * n_src, 2 * n_src, sg2,
                    1.0, h_out + (int64_t)e * m * 2 * n_src, 2 * n_src, sH, n_merges));
      if (BD_inv)
        HPS_TRY(dgemm(st, m, 2 * n, 2 * m, 1.0, Ablk, 2 * m, sB, Di2 + (int64_t)2 * slot * m * n2, n2, sS2,
                      j == 0 ? 0.0 : 1.0, BD_inv + (int64_t)e * m * n2, n2, sR, n_merges));
    }
  return 0;
}

int up_gather_quad_iti(cudaStream_t st, int n_nodes, int m, int n_src, const double* h_in, double* h_int, double* h_ext,
                       int ext_shift, const int* pos8) {
  if (n_nodes <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  IntPos ip;
  for (int u = 0; u < 8; ++u) ip.pos[u] = pos8 ? pos8[u] : u;
  const int64_t total = (int64_t)16 * m * n_src;
  dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 1024), n_nodes);
  up_gather_iti_kernel<<<grid, 256, 0, st>>>(iti_topo(), m, n_src, reinterpret_cast<const double2*>(h_in),
                                             reinterpret_cast<double2*>(h_int), reinterpret_cast<double2*>(h_ext), ext_shift, ip);
  HPS_LAUNCH_CHECK("up_gather_iti_kernel");
  return 0;
}

// ws: n_nodes * (16m * 2 n_src + 8m * 2 n_src) doubles
int down_quad_iti_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                        const double* gt, double* g_children, void* ws) {
  if (n_nodes <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  if (n_nodes > 65535) return fail_arg(2, "n_nodes per call is limited to 65535");
  const ItiTopo& tp = iti_topo();
  const int n = 8 * m, n2 = 2 * n, w = 2 * n_src;
  double* G2 = static_cast<double*>(ws);
  double* g_int = G2 + (int64_t)n_nodes * n2 * w;
  HPS_TRY(complex_expand(st, n_nodes, n, n_src, g_ext, (int64_t)n * n_src, G2, (int64_t)n2 * w));
  HPS_TRY(dgemm_affine(st, n, w, n2, S, n2, (int64_t)n * n2, G2, w, (int64_t)n2 * w, gt, w, (int64_t)n * w, g_int, w,
                       (int64_t)n * w, n_nodes));
  return down_scatter(tp.base, st, n_nodes, m, w, g_ext, g_int, g_children);
}

// u = Y g + v on every leaf, complex128.  ws: n_leaves * 2 n_g * 2 n_src doubles
int leaf_apply_complex(cudaStream_t st, int n_leaves, int n_c, int n_g, int n_src, const double* Y, const double* g,
                       const double* v, double* u, void* ws) {
  if (n_leaves <= 0 || n_c <= 0 || n_g <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  const int w = 2 * n_src;
  double* G2 = static_cast<double*>(ws);
  HPS_TRY(complex_expand(st, n_leaves, n_g, n_src, g, (int64_t)n_g * n_src, G2, (int64_t)2 * n_g * w));
  return dgemm_affine(st, n_c, w, 2 * n_g, Y, 2 * n_g, (int64_t)n_c * 2 * n_g, G2, w, (int64_t)2 * n_g * w, v, w,
                      (int64_t)n_c * w, u, w, (int64_t)n_c * w, n_leaves);
}

size_t merge_oct_ws_bytes(int n_merges, int m) { return merge_ws_bytes(oct_topo(), n_merges, m); }
size_t merge_quad_ws_bytes(int n_merges, int m) { return merge_ws_bytes(quad_topo(), n_merges, m); }

int merge_oct_level(cudaStream_t st, int n_merges, int m, int n_src, const double* T_in, const double* h_in, double* S,
                    double* gt, double* T_out, double* h_out, int want_T, void* ws, size_t ws_bytes, int* info) {
  return merge_level(oct_topo(), st, n_merges, m, n_src, T_in, h_in, S, gt, T_out, h_out, want_T, ws, ws_bytes, info);
}
int merge_quad_level(cudaStream_t st, int n_merges, int m, int n_src, const double* T_in, const double* h_in,
                     double* S, double* gt, double* T_out, double* h_out, int want_T, void* ws, size_t ws_bytes,
                     int* info) {
  return merge_level(quad_topo(), st, n_merges, m, n_src, T_in, h_in, S, gt, T_out, h_out, want_T, ws, ws_bytes, info);
}
int merge_quad_level_nosource(cudaStream_t st, int n_merges, int m, const double* T_in, double* S, double* T_out,
                              double* D_inv, double* BD_inv, double* h_zero, double* scratch_h, void* ws,
                              size_t ws_bytes, int* info) {
  // h_zero: n_merges*4*4m zeros (children's h), scratch_h: n_merges*(4m + 8m) doubles for g~/h_out (discarded)
  const int n_int = 4 * m;
  return merge_level(quad_topo(), st, n_merges, m, 1, T_in, h_zero, S, scratch_h, T_out,
                     scratch_h + (int64_t)n_merges * n_int, 1, ws, ws_bytes, info, D_inv, BD_inv);
}

int up_gather_quad(cudaStream_t st, int n_nodes, int m, int n_src, int is_complex, const double* h_in, double* h_int,
                   double* h_ext, int ext_shift) {
  if (n_nodes <= 0 || m <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  const Topo& tp = quad_topo();
  const int64_t total = (int64_t)(12 * m) * n_src;
  dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 1024), n_nodes);
  if (is_complex)
    up_gather_dtn_kernel<double2><<<grid, 256, 0, st>>>(tp, m, n_src, reinterpret_cast<const double2*>(h_in),
                                                       reinterpret_cast<double2*>(h_int), reinterpret_cast<double2*>(h_ext), ext_shift);
  else
    up_gather_dtn_kernel<double><<<grid, 256, 0, st>>>(tp, m, n_src, h_in, h_int, h_ext, ext_shift);
  HPS_LAUNCH_CHECK("up_gather_dtn_kernel");
  return 0;
}

int down_oct_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                   const double* gt, double* g_children, void* ws) {
  return down_level(oct_topo(), st, n_nodes, m, n_src, S, g_ext, gt, g_children, ws);
}
int down_quad_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                    const double* gt, double* g_children, void* ws) {
  return down_level(quad_topo(), st, n_nodes, m, n_src, S, g_ext, gt, g_children, ws);
}

}  // namespace hps
