// History-ring companions of the env kernel: rows of the iteration out of the ring, and the
// hand-over of the last row's window to the next iteration.
//
// Reference being replaced: the re-stacking of all frames on every step (humanoid_env.py:880-887: `torch.stack` of the 15
// deque entries + reshape, `torch.cat` of the 3 critic frames) and the ten index gathers of mini_batch_generator
// (rollout_storage.py:176-189) for the two observation tensors.  The ring keeps ONE copy of every frame; a stacked row is
// the window [n, r .. r+frames-1, :] of the env-major archive, with the frames older than the env's last reset read as zero
// (reset_idx zeroes the deque entries, :1159-1163).
#include "hg_common.cuh"

namespace hg {

// One warp per output row.  fp32 output (the reference's layout, packed): the window is `frames * width` contiguous floats of
// the fp32 archive.  bf16 output (the minibatch operand, frame pitch `pitch16`): the window is `frames * pitch16 / 8`
// contiguous 16-byte chunks of the bf16 archive, copied chunk for chunk (a whole frame is valid or zero).
template <bool BF16>
__global__ void __launch_bounds__(256) ring_gather_kernel(const HgHistoryRing R, int num_envs, const int* __restrict__ first_valid,
                                                          const long long* __restrict__ idx, long long rows_out, void* __restrict__ dst,
                                                          long long ld_dst, float* __restrict__ lv_out, int lv_col) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < rows_out; i += warps) {
    const long long flat = idx ? __ldg(idx + i) : i;
    const int r = (int)(flat / num_envs), n = (int)(flat - (long long)r * num_envs);
    const int fv = __ldg(first_valid + flat);                  // frames whose time (r - frames + 1 + f) is older read as zero
    const float* src = R.arch32 + ((size_t)n * R.slots + r) * R.width;
    const int f0 = max(0, fv - (r - R.frames + 1));            // first valid frame of the window
    if (dst == nullptr) {
      // only the lin-vel targets are wanted
    } else if constexpr (BF16) {
      const int cpf = R.pitch16 >> 3;                          // 16-byte chunks per frame
      const uint4* s16 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(R.arch16) + ((size_t)n * R.slots + r) * R.pitch16);
      uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + i * ld_dst);
      const int chunks = R.frames * cpf;
      constexpr int U = 4;                                     // loads in flight per lane (a 15 x 48 window is 90 chunks: one pass)
      for (int c0 = 0; c0 < chunks; c0 += 32 * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + lane + 32 * u;
          v[u] = make_uint4(0, 0, 0, 0);
          if (c < chunks && c / cpf >= f0) v[u] = __ldg(s16 + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int c = c0 + lane + 32 * u;
          if (c < chunks) out[c] = v[u];
        }
      }
    } else {
      float* out = reinterpret_cast<float*>(dst) + i * ld_dst;
      for (int j = lane; j < R.frames * R.width; j += 32) out[j] = (j / R.width >= f0) ? __ldg(src + j) : 0.f;
    }
    if (lv_out && lane < 3) {
      const int j = lv_col + lane;
      lv_out[i * 4 + lane] = (j / R.width >= f0) ? __ldg(src + j) : 0.f;
    }
  }
}

// Targets only (dst == nullptr): one THREAD per row, three floats each (a warp per row left 29 lanes idle: 41 us for 245,760 rows).
__global__ void __launch_bounds__(256) ring_targets_kernel(const HgHistoryRing R, int num_envs, const int* __restrict__ first_valid,
                                                           const long long* __restrict__ idx, long long rows_out,
                                                           float* __restrict__ lv_out, int lv_col) {
  pdl_trigger();
  pdl_wait();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows_out; i += stride) {
    const long long flat = idx ? __ldg(idx + i) : i;
    const int r = (int)(flat / num_envs), n = (int)(flat - (long long)r * num_envs);
    const int f0 = max(0, __ldg(first_valid + flat) - (r - R.frames + 1));
    const float* src = R.arch32 + ((size_t)n * R.slots + r) * R.width;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((lv_col + 0) / R.width >= f0) o.x = __ldg(src + lv_col);
    if ((lv_col + 1) / R.width >= f0) o.y = __ldg(src + lv_col + 1);
    if ((lv_col + 2) / R.width >= f0) o.z = __ldg(src + lv_col + 2);
    *reinterpret_cast<float4*>(lv_out + i * 4) = o;
  }
}

// Hand-over without overlap (from_row >= frames, e.g. 60 >= 15): source and destination windows are disjoint, so the copy is
// a plain grid-stride loop over [ring][env][frames * width] floats -- the register-staged kernel below (one warp per env,
// 96 values per lane) took 84 us for 33 MB.
__global__ void __launch_bounds__(256) ring_rollover_disjoint_kernel(const HgHistoryRing R0, const HgHistoryRing R1, int num_rings, int num_envs,
                                                                     int from_row, int* __restrict__ hist_start, int* __restrict__ first_valid_row0) {
  pdl_trigger();
  pdl_wait();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (int ring = 0; ring < num_rings; ++ring) {
    const HgHistoryRing& R = ring == 0 ? R0 : R1;
    const int count = R.frames * R.width, count16 = R.frames * (R.pitch16 >> 1);     // floats / bf16 pairs per window
    for (long long i = tid; i < (long long)num_envs * count; i += stride) {
      const int n = (int)(i / count), j = (int)(i - (long long)n * count);
      const size_t base = (size_t)n * R.slots * R.width;
      R.live32[base + j] = R.live32[base + (size_t)from_row * R.width + j];
      R.arch32[base + j] = R.arch32[base + (size_t)from_row * R.width + j];
    }
    const uint32_t* l16 = reinterpret_cast<const uint32_t*>(R.live16);
    const uint32_t* a16 = reinterpret_cast<const uint32_t*>(R.arch16);
    const int pairs = R.pitch16 >> 1;
    for (long long i = tid; i < (long long)num_envs * count16; i += stride) {
      const int n = (int)(i / count16), j = (int)(i - (long long)n * count16);
      const size_t base = (size_t)n * R.slots * pairs;
      reinterpret_cast<uint32_t*>(R.live16)[base + j] = l16[base + (size_t)from_row * pairs + j];
      reinterpret_cast<uint32_t*>(R.arch16)[base + j] = a16[base + (size_t)from_row * pairs + j];
    }
  }
  for (long long n = tid; n < num_envs; n += stride) {
    const int hs = max(hist_start[n] - from_row, -(1 << 24));
    hist_start[n] = hs;
    first_valid_row0[n] = hs;
  }
}

// One warp per (ring, env): the whole window of `from_row` is read into registers before the first store (source and
// destination slots overlap when from_row < frames - 1).
constexpr int ROLL_MAX = 48;      // floats per lane: frames * width <= 32 * 48 = 1536 (15 x 97 = 1455 for the 18-DOF layout)
__global__ void __launch_bounds__(128) ring_rollover_kernel(const HgHistoryRing R0, const HgHistoryRing R1, int num_rings, int num_envs,
                                                            int from_row, int* __restrict__ hist_start, int* __restrict__ first_valid_row0) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= num_rings * num_envs) return;
  const int ring = w / num_envs, n = w - ring * num_envs;
  const HgHistoryRing& R = ring == 0 ? R0 : R1;
  const size_t base = (size_t)n * R.slots;
  const int count = R.frames * R.width;
  float live[ROLL_MAX], arch[ROLL_MAX];
#pragma unroll
  for (int u = 0; u < ROLL_MAX; ++u) {
    const int j = lane + 32 * u;
    live[u] = j < count ? R.live32[(base + from_row) * R.width + j] : 0.f;
    arch[u] = j < count ? R.arch32[(base + from_row) * R.width + j] : 0.f;
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < ROLL_MAX; ++u) {
    const int j = lane + 32 * u;
    if (j < count) {
      R.live32[base * R.width + j] = live[u];
      R.arch32[base * R.width + j] = arch[u];
    }
  }
  // bf16 copies of the two windows, re-derived from the fp32 values just moved (same rounding as the env kernel)
  __nv_bfloat16* l16 = reinterpret_cast<__nv_bfloat16*>(R.live16) + base * R.pitch16;
  __nv_bfloat16* a16 = reinterpret_cast<__nv_bfloat16*>(R.arch16) + base * R.pitch16;
#pragma unroll
  for (int u = 0; u < ROLL_MAX; ++u) {
    const int j = lane + 32 * u;
    if (j < count) {
      const int f = j / R.width, c = j - f * R.width;
      l16[f * R.pitch16 + c] = __float2bfloat16_rn(live[u]);
      a16[f * R.pitch16 + c] = __float2bfloat16_rn(arch[u]);
    }
  }
  if (ring == 0 && lane == 0) {
    const int hs = max(hist_start[n] - from_row, -(1 << 24));
    hist_start[n] = hs;
    first_valid_row0[n] = hs;
  }
}

static int check_ring(const HgHistoryRing* R, const char* who) {
  HG_CHECK(R && R->live32 && R->live16 && R->arch32 && R->arch16, "%s: null ring buffer", who);
  HG_CHECK(R->width >= 1 && R->frames >= 1 && R->pitch16 >= R->width && R->pitch16 % 8 == 0 && R->slots >= R->frames,
           "%s: bad ring geometry (width %d, pitch16 %d, frames %d, slots %d)", who, R->width, R->pitch16, R->frames, R->slots);
  return HG_OK;
}

}  // namespace hg

extern "C" int hg_ring_gather(const HgHistoryRing* ring, int num_envs, const int32_t* first_valid, const int64_t* idx,
                              int64_t rows_out, void* dst, int64_t ld_dst, int dst_bf16, float* lv_out, int lv_col, void* stream) {
  using namespace hg;
  int rc = check_ring(ring, "hg_ring_gather");
  if (rc != HG_OK) return rc;
  HG_CHECK(num_envs > 0 && first_valid && (dst || lv_out) && rows_out > 0, "hg_ring_gather: bad arguments");
  HG_CHECK(!dst || (ld_dst >= (int64_t)ring->frames * (dst_bf16 ? ring->pitch16 : ring->width) && (!dst_bf16 || ld_dst % 8 == 0)),
           "hg_ring_gather: ld_dst %lld too small for %d frames (bf16 rows: a multiple of 8)", (long long)ld_dst, ring->frames);
  HG_CHECK(!dst || !dst_bf16 || ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(ring->arch16)) & 15) == 0,
           "hg_ring_gather: bf16 rows and the bf16 archive must be 16-byte aligned");
  HG_CHECK(!lv_out || (lv_col >= 0 && lv_col + 3 <= ring->frames * ring->width), "hg_ring_gather: lv_col outside the row");
  long long blocks = (rows_out + 7) / 8;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long* idx_ll = reinterpret_cast<const long long*>(idx);
  if (dst == nullptr) {
    HG_CHECK((reinterpret_cast<uintptr_t>(lv_out) & 15) == 0, "hg_ring_gather: lv_out must be 16-byte aligned");
    long long tb = (rows_out + 255) / 256;
    if (tb > (long long)num_sms() * 8) tb = (long long)num_sms() * 8;
    HG_CUDA(launch_kernel(ring_targets_kernel, dim3((unsigned)tb), dim3(256), 0, st, 1, *ring, num_envs, first_valid, idx_ll,
                          (long long)rows_out, lv_out, lv_col));
    HG_LAUNCH_CHECK();
    count_launch();
    return HG_OK;
  }
  if (dst_bf16)
    HG_CUDA(launch_kernel(ring_gather_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, st, 1, *ring, num_envs, first_valid, idx_ll,
                          (long long)rows_out, dst, (long long)ld_dst, lv_out, lv_col));
  else
    HG_CUDA(launch_kernel(ring_gather_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, 1, *ring, num_envs, first_valid, idx_ll,
                          (long long)rows_out, dst, (long long)ld_dst, lv_out, lv_col));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_ring_rollover(const HgHistoryRing* rings, int num_rings, int num_envs, int from_row, int32_t* hist_start,
                                int32_t* first_valid_row0, void* stream) {
  using namespace hg;
  HG_CHECK(rings && num_rings >= 1 && num_rings <= 2 && num_envs > 0 && hist_start && first_valid_row0, "hg_ring_rollover: bad arguments");
  for (int i = 0; i < num_rings; ++i) {
    int rc = check_ring(rings + i, "hg_ring_rollover");
    if (rc != HG_OK) return rc;
    HG_CHECK(rings[i].frames * rings[i].width <= 32 * ROLL_MAX, "hg_ring_rollover: window of %d floats too wide", rings[i].frames * rings[i].width);
    HG_CHECK(from_row >= 0 && from_row + rings[i].frames <= rings[i].slots, "hg_ring_rollover: from_row %d outside the ring", from_row);
  }
  if (from_row == 0) return HG_OK;
  bool disjoint = true;
  for (int i = 0; i < num_rings; ++i) disjoint = disjoint && from_row >= rings[i].frames;
  if (disjoint) {
    HG_CUDA(launch_kernel(ring_rollover_disjoint_kernel, dim3(num_sms() * 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1,
                          rings[0], rings[num_rings - 1], num_rings, num_envs, from_row, hist_start, first_valid_row0));
    HG_LAUNCH_CHECK();
    count_launch();
    return HG_OK;
  }
  const int warps = num_rings * num_envs;
  HG_CUDA(launch_kernel(ring_rollover_kernel, dim3((warps + 3) / 4), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 1, rings[0],
                        rings[num_rings - 1], num_rings, num_envs, from_row, hist_start, first_valid_row0));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
