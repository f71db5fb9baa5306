// K7: cyclic-shift roll, window partition / reverse, shift mask, position index.
// Pure index work -- bit-exact against models/layers.py:187-216, 224-293, 371-411.
#include "common.cuh"

namespace pangu {

// One thread per 16-byte vector of a window row; consecutive threads -> consecutive vectors, so
// both the gather (partition) and the scatter (reverse) move whole token rows coalesced.
template <bool kReverse>
__global__ void __launch_bounds__(256)
window_move_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, WinGeom g, int roll,
                   int vecs_per_row, long long total_vecs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total_vecs) return;
  const long long row = gid / vecs_per_row;
  const int v = (int)(gid - row * vecs_per_row);
  const int k = (int)(row % kWinTokens);
  const long long lt = row / kWinTokens;
  const int t = (int)(lt % g.T);
  const int l = (int)(lt / g.T);
  const long long n = window_source(g, l, t, k, roll);
  if (!kReverse) {
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (n >= 0) val = __ldg(src + n * vecs_per_row + v);
    dst[gid] = val;
  } else {
    if (n >= 0) dst[n * vecs_per_row + v] = __ldg(src + gid);
  }
}

__global__ void window_source_index_kernel(long long* __restrict__ idx, WinGeom g, int roll,
                                           long long total) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int k = (int)(gid % kWinTokens);
  const long long lt = gid / kWinTokens;
  idx[gid] = window_source(g, (int)(lt / g.T), (int)(lt % g.T), k, roll);
}

__global__ void shift_mask_kernel(float* __restrict__ mask, WinGeom g, long long total) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int j = (int)(gid % kWinTokens);
  const long long ti = gid / kWinTokens;
  const int i = (int)(ti % kWinTokens);
  const int t = (int)(ti / kWinTokens);
  // models/layers.py:212: mask_windows.unsqueeze(2) - mask_windows.unsqueeze(3), then != 0 -> -100
  mask[gid] = shift_region_reference(g, t, i) != shift_region_reference(g, t, j) ? kMaskValue : 0.0f;
}

__global__ void position_index_kernel(long long* __restrict__ idx) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= kWinTokens * kWinTokens) return;
  const int i = gid / kWinTokens, j = gid - i * kWinTokens;
  const int zi = i / 72, hi = (i / 12) % 6, wi = i % 12;
  const int zj = j / 72, hj = (j / 12) % 6, wj = j % 12;
  // coords_1 - coords_2 with coords_2 = (-2*zj, -6*hj, wj); then the shifts of layers.py:402-405
  idx[gid] = (long long)(zi + 2 * zj) * (23 * 36) + (hi + 6 * hj) * 23 + (wi - wj + 11);
}

__device__ __forceinline__ int bias_table_index(int i, int j) {
  const int zi = i / 72, hi = (i / 12) % 6, wi = i % 12;
  const int zj = j / 72, hj = (j / 12) % 6, wj = j % 12;
  return (zi + 2 * zj) * (23 * 36) + (hi + 6 * hj) * 23 + (wi - wj + 11);
}

// Compact Earth-specific bias (the parameterisation of the paper, kept in the reference as commented code,
// models/layers.py:355,442-449): full[t, h, i, j] = table[position_index[i*144+j], t, h].  One thread per output element,
// writes coalesced; the [3312, T, heads] table (<= 10 MB) stays in L2.
__global__ void __launch_bounds__(256)
bias_table_expand_kernel(const float* __restrict__ table, float* __restrict__ full, int T, int heads) {
  const long long total = (long long)T * heads * kWinTokens * kWinTokens;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int ij = (int)(g % (kWinTokens * kWinTokens));
    const int th = (int)(g / (kWinTokens * kWinTokens));           // t * heads + h
    full[g] = __ldg(table + (long long)bias_table_index(ij / kWinTokens, ij % kWinTokens) * T * heads + th);
  }
}

// Adjoint: d_table[idx, t, h] += sum over the (i, j) pairs that share idx.  (zi, zj) and (hi, hj) are determined by idx; only
// the longitude offset d = wi - wj is shared, by 12 - |d| pairs: a deterministic gather-reduce, no atomics.
__global__ void __launch_bounds__(256)
bias_table_reduce_kernel(const float* __restrict__ d_full, float* __restrict__ d_table, int T, int heads) {
  const long long total = 3312LL * T * heads;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int th = (int)(g % ((long long)T * heads));
    const int idx = (int)(g / ((long long)T * heads));
    const int zz = idx / (23 * 36), r = idx - zz * (23 * 36), hh = r / 23, d = r - hh * 23 - 11;
    const int zi = zz & 1, zj = zz >> 1, hi = hh % 6, hj = hh / 6;
    const float* src = d_full + (long long)th * kWinTokens * kWinTokens;
    float acc = 0.f;
    for (int wj = max(0, -d); wj < min(12, 12 - d); ++wj) {
      const int wi = wj + d;
      acc += __ldg(src + (zi * 72 + hi * 12 + wi) * kWinTokens + (zj * 72 + hj * 12 + wj));
    }
    d_table[g] += acc;
  }
}

static int window_move(const void* src, void* dst, const pangu_geom* gg, int roll, int elem_bytes,
                       void* stream, bool reverse) {
  WinGeom g;
  if (!make_geom(gg, g) || !src || !dst) { set_error("window_move: bad geometry or null pointer"); return PANGU_ERR_BAD_ARG; }
  const long long row_bytes = (long long)g.C * elem_bytes;
  if ((elem_bytes != 2 && elem_bytes != 4) || row_bytes % 16) {
    set_error("window_move: C*elem_bytes=%lld must be a multiple of 16", row_bytes);
    return PANGU_ERR_BAD_ARG;
  }
  const int vpr = (int)(row_bytes / 16);
  const long long total = (long long)g.nLon * g.T * kWinTokens * vpr;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  if (reverse)
    window_move_kernel<true><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(
        (const uint4*)src, (uint4*)dst, g, roll, vpr, total);
  else
    window_move_kernel<false><<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(
        (const uint4*)src, (uint4*)dst, g, roll, vpr, total);
  return check_launch(reverse ? "window_reverse" : "window_partition");
}

}  // namespace pangu

using namespace pangu;

extern "C" int pangu_window_partition(const void* x, void* win, const pangu_geom* g, int roll,
                                      int elem_bytes, void* stream) {
  return window_move(x, win, g, roll, elem_bytes, stream, false);
}

extern "C" int pangu_window_reverse(const void* win, void* x, const pangu_geom* g, int roll,
                                    int elem_bytes, void* stream) {
  return window_move(win, x, g, roll, elem_bytes, stream, true);
}

extern "C" int pangu_window_source_index(int64_t* idx, const pangu_geom* gg, int roll, void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !idx) { set_error("window_source_index: bad argument"); return PANGU_ERR_BAD_ARG; }
  const long long total = (long long)g.nLon * g.T * kWinTokens;
  window_source_index_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      (long long*)idx, g, roll, total);
  return check_launch("window_source_index");
}

extern "C" int pangu_shift_mask(float* mask, const pangu_geom* gg, void* stream) {
  WinGeom g;
  if (!make_geom(gg, g) || !mask) { set_error("shift_mask: bad argument"); return PANGU_ERR_BAD_ARG; }
  const long long total = (long long)g.T * kWinTokens * kWinTokens;
  shift_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(mask, g, total);
  return check_launch("shift_mask");
}

extern "C" int pangu_position_index(int64_t* idx, void* stream) {
  if (!idx) { set_error("position_index: null"); return PANGU_ERR_BAD_ARG; }
  position_index_kernel<<<(kWinTokens * kWinTokens + 255) / 256, 256, 0, as_stream(stream)>>>((long long*)idx);
  return check_launch("position_index");
}

extern "C" int pangu_bias_table_expand(const float* table, float* full, int32_t T, int32_t heads, void* stream) {
  if (!table || !full || T <= 0 || heads <= 0) { set_error("bias_table_expand: bad argument"); return PANGU_ERR_BAD_ARG; }
  bias_table_expand_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(table, full, T, heads);
  return check_launch("bias_table_expand");
}

extern "C" int pangu_bias_table_reduce(const float* d_full, float* d_table, int32_t T, int32_t heads, void* stream) {
  if (!d_full || !d_table || T <= 0 || heads <= 0) { set_error("bias_table_reduce: bad argument"); return PANGU_ERR_BAD_ARG; }
  bias_table_reduce_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(d_full, d_table, T, heads);
  return check_launch("bias_table_reduce");
}
