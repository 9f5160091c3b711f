// Small HBM-bound data-movement kernels of the learner: obs ingest (fp32 -> storage slot + bf16 shadow),
// minibatch row gather, weight cast, reward bootstrap.
#include "hg_common.cuh"

namespace hg {

// One warp per row.  Row r: read `cols` fp32 (coalesced 128 B per warp), optionally copy them to the
// fp32 storage slot, and write the bf16 shadow row (ld16 elements, zero padded) as 4-byte bf16x2 stores.
__global__ void ingest_rows_kernel(const float* __restrict__ src, long long src_ld, int rows, int cols,
                                   float* __restrict__ dst, long long dst_ld, __nv_bfloat16* __restrict__ dst16,
                                   long long ld16) {
  pdl_trigger();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const float* s = src + (long long)r * src_ld;
    float* d = dst ? dst + (long long)r * dst_ld : nullptr;
    uint32_t* d16 = reinterpret_cast<uint32_t*>(dst16 + (long long)r * ld16);
    for (int c = 2 * lane; c < ld16; c += 64) {
      const float a = c < cols ? s[c] : 0.f;
      const float b = c + 1 < cols ? s[c + 1] : 0.f;
      if (d) {
        if (c < cols) d[c] = a;
        if (c + 1 < cols) d[c + 1] = b;
      }
      d16[c >> 1] = pack_bf16x2(a, b);
    }
  }
}

// One warp per destination row, 16-byte chunks.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, int chunks_per_row, const long long* __restrict__ idx,
                                   int rows, uint4* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    const uint4* s = src + idx[r] * chunks_per_row;
    uint4* d = dst + (long long)r * chunks_per_row;
    for (int c = lane; c < chunks_per_row; c += 32) d[c] = __ldg(s + c);
  }
}

__global__ void cast_weight_kernel(const float* __restrict__ src, int rows, int cols, __nv_bfloat16* __restrict__ dst,
                                   long long ld, int src_frame, int dst_frame) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i % ld);
    int sc = c;                       // source column of destination column c (-1: padding)
    if (src_frame) {
      const int f = c / dst_frame, k = c - f * dst_frame;
      sc = k < src_frame ? f * src_frame + k : -1;
    }
    dst[i] = __float2bfloat16_rn(sc >= 0 && sc < cols ? src[(long long)r * cols + sc] : 0.f);
  }
}

// rewards.clone() += gamma * (values * time_outs)   (reference ppo.py:129-133), dones -> uint8 (rollout_storage.py:94)
__global__ void store_rewards_dones_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                           const uint8_t* __restrict__ time_outs, const uint8_t* __restrict__ dones,
                                           float gamma, int n, float* __restrict__ out_r, uint8_t* __restrict__ out_d) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r = rewards[i];
  if (time_outs) r = __fadd_rn(r, __fmul_rn(gamma, __fmul_rn(values[i], time_outs[i] ? 1.f : 0.f)));
  out_r[i] = r;
  out_d[i] = dones[i] ? 1 : 0;
}

static inline int row_grid(int rows, int warps_per_block) {
  long long b = ((long long)rows + warps_per_block - 1) / warps_per_block;
  const long long cap = (long long)num_sms() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace hg

extern "C" int hg_ingest_rows(const float* src, int64_t src_ld, int rows, int cols, float* dst_f32, int64_t dst_ld,
                              void* dst_bf16, int64_t ld16, void* stream) {
  using namespace hg;
  HG_CHECK(src && dst_bf16 && rows > 0 && cols > 0, "hg_ingest_rows: bad arguments");
  HG_CHECK(ld16 >= cols && ld16 % 8 == 0, "hg_ingest_rows: ld16 must be >= cols and a multiple of 8");
  HG_CHECK((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0, "hg_ingest_rows: bf16 destination alignment");
  HG_CUDA(launch_kernel(ingest_rows_kernel, dim3(row_grid(rows, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, src, src_ld, rows, cols, dst_f32, dst_ld, reinterpret_cast<__nv_bfloat16*>(dst_bf16), ld16));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_gather_rows_bf16(const void* src, int64_t ld, const int64_t* idx, int rows, void* dst, void* stream) {
  using namespace hg;
  HG_CHECK(src && idx && dst && rows > 0, "hg_gather_rows_bf16: bad arguments");
  HG_CHECK(ld % 8 == 0, "hg_gather_rows_bf16: ld must be a multiple of 8");
  HG_CHECK(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "hg_gather_rows_bf16: alignment");
  HG_CUDA(launch_kernel(gather_rows_kernel, dim3(row_grid(rows, 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, reinterpret_cast<const uint4*>(src), (int)(ld / 8), reinterpret_cast<const long long*>(idx), rows, reinterpret_cast<uint4*>(dst)));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_cast_weight_bf16(const float* src, int rows, int cols, void* dst, int64_t ld, int src_frame, int dst_frame, void* stream) {
  using namespace hg;
  HG_CHECK(src && dst && rows > 0 && cols > 0 && ld >= cols, "hg_cast_weight_bf16: bad arguments");
  HG_CHECK((src_frame == 0 && dst_frame == 0) || (src_frame >= 1 && dst_frame >= src_frame && cols % src_frame == 0 &&
                                                  ld >= (int64_t)(cols / src_frame) * dst_frame),
           "hg_cast_weight_bf16: bad frame remap %d -> %d for %d columns (ld %lld)", src_frame, dst_frame, cols, (long long)ld);
  const long long total = (long long)rows * ld;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  HG_CUDA(launch_kernel(cast_weight_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, src, rows, cols, reinterpret_cast<__nv_bfloat16*>(dst), (long long)ld, src_frame, dst_frame));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

extern "C" int hg_store_rewards_dones(const float* rewards, const float* values, const uint8_t* time_outs,
                                      const uint8_t* dones, float gamma, int n, float* out_rewards, uint8_t* out_dones,
                                      void* stream) {
  using namespace hg;
  HG_CHECK(rewards && values && dones && out_rewards && out_dones && n > 0, "hg_store_rewards_dones: bad arguments");
  HG_CUDA(launch_kernel(store_rewards_dones_kernel, dim3((n + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 1, rewards, values, time_outs, dones, gamma, n, out_rewards, out_dones));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}
