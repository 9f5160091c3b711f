// Shared device/host helpers for the sm_100a kernels of the humanoid-gym PPO hot path.
// Raw PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hgb200.h"

namespace hg {

// last error text, readable through hg_last_error()
void set_error(const char* fmt, ...);

#define HG_CHECK(cond, ...)                                   \
  do {                                                        \
    if (!(cond)) {                                            \
      ::hg::set_error(__VA_ARGS__);                           \
      return HG_ERR_INVALID;                                  \
    }                                                         \
  } while (0)

#define HG_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::hg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return HG_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define HG_LAUNCH_CHECK() HG_CUDA(cudaGetLastError())

int num_sms();
void count_launch(int n = 1);
bool pdl_enabled();   // programmatic dependent launch on (default) unless HG_PDL=0

// Every kernel of the library is launched through here.  With programmatic dependent launch the next kernel of the
// stream becomes resident while this one drains: its on-chip prologue (barrier init, TMEM allocation, descriptor
// prefetch) overlaps the tail of its predecessor, and `pdl_wait()` (griddepcontrol.wait) orders every global access
// after the predecessor's completion + flush.  Contract: each kernel calls pdl_wait() BEFORE its first global-memory
// access.  Measured on B200 (4096 envs x 60, one PPO iteration): no PDL 10.69 ms, PDL with the implicit trigger at
// grid exit 9.63 ms, PDL with an explicit early trigger (griddepcontrol.launch_dependents at kernel entry) 11.30 ms
// -- early-resident dependents take SM slots from the kernel that is still running -- so pdl_trigger() is a no-op
// unless HG_PDL_EARLY_TRIGGER is defined.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void pdl_trigger() {
#ifdef HG_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// LATE trigger: the dependent grid is launched once EVERY block of this grid has executed this (or exited), so a kernel
// that calls it when its blocks are past their main work -- and therefore all resident -- lets the ~2 us launch latency
// of its successor (measured with tools/trace_rollout.py: 1.8-2.0 us between the last block's exit and the first
// instruction of the next kernel) overlap its own tail instead of following it.  The successor still orders its global
// accesses behind this grid's completion with pdl_wait().  One executing thread per block is enough.
// Measured (tools/phase_times.py, variants with / without): rollout 3.40 vs 3.38 ms, update 5.47 vs 5.36 ms -- the wait of
// the successor for this grid's completion + flush (~1.3 us) is there either way and what the trigger saves in launch
// latency it loses elsewhere, so the trigger is OFF unless HG_LATE_TRIGGER is defined.
__device__ __forceinline__ void pdl_launch_dependents() {
#ifdef HG_LATE_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware (up to ~the hint) until the phase completes instead of
// spinning -- the spin loop of waiting epilogue warps (try_wait + clock64 + compare + branch) was 7 % of all instructions
// issued by the dX launches, taken from the sibling warps that still had work on the same scheduler.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug (wrong expect_tx, lost arrive) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: surfaces as a launch failure on the host
  }
}

// ---- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane_base + t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- device-side timeline (development builds only: tools/build_variant.sh NAME -DHG_TRACE, tools/trace_rollout.py)
// trace(tag, sub) stores %globaltimer into a fixed slot of a table: [kernel = tag >> 8 (1..3)][point = tag & 63][sub 0..3]
// [block 0..511] -- plain stores, no atomics (a shared record counter serialised the blocks and distorted the very
// timeline it was measuring), so a table holds ONE launch of each kernel: the tool traces one rollout step at a time.
// Each translation unit keeps its own table pointer (no relocatable device code); hg_trace_set() sets them all.
#ifdef HG_TRACE
static __device__ unsigned long long* g_trace_buf = nullptr;
constexpr int TRACE_BLOCKS = 512;
__device__ __forceinline__ void trace(int tag, int sub = 0) {
  unsigned long long* b = g_trace_buf;
  if (!b || blockIdx.x >= TRACE_BLOCKS) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  b[((size_t)((((tag >> 8) & 3) * 64 + (tag & 63)) * 4 + (sub & 3))) * TRACE_BLOCKS + blockIdx.x] = t;
}
#define HG_TRACE_SETTER(name) \
  int name(void* p) { return (int)cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)); }
#else
__device__ __forceinline__ void trace(int, int = 0) {}
#define HG_TRACE_SETTER(name)
#endif

// ---- small math helpers
__device__ __forceinline__ float elu1(float z) { return z > 0.f ? z : expm1f(z); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float LOG_SQRT_2PI = 0.9189385332046727f;  // log(sqrt(2*pi)) as torch.distributions.Normal uses it

// ------------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  const float u2 = b * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}
// standard normal number `j` of row `row` for (seed, counter)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t counter, uint32_t row, uint32_t j) {
  uint4 r = philox4x32_10(make_uint4(row, j >> 2, (uint32_t)counter, (uint32_t)(counter >> 32)),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 n01 = box_muller(r.x, r.y), n23 = box_muller(r.z, r.w);
  const int k = j & 3;
  return k == 0 ? n01.x : k == 1 ? n01.y : k == 2 ? n23.x : n23.y;
}


}  // namespace hg
