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
