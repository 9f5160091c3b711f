// Grouped bf16 GEMM on 5th-generation tensor cores (tcgen05.mma, TMEM accumulators, TMA loads).
//
//   C[M,N] = A[M,K] * B[N,K]^T   (fp32 accumulate), up to HG_MAX_GROUP problems per launch.
//
// One persistent CTA per SM, 20 warps:
//   warp 0     TMA producer (one lane): cp.async.bulk.tensor 2D boxes -> SWIZZLE_128B smem ring
//   warp 1     MMA issuer  (one lane): tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16 per instruction
//   warp 2     TMEM allocator (2*BLOCK_N columns: two accumulator buffers so the epilogue of tile i
//              overlaps the main loop of tile i+1)
//   warps 4-19 epilogue (four per scheduler so they hide each other's latency): warp w reads TMEM lanes
//              32*(w%4).. and the column slice (w-4)/4 with tcgen05.ld 32x32b -> registers -> fused epilogue -> global
//
// Operands may be K-major (row-major [rows,K]) or MN-major (row-major [K,rows]); the latter lets the
// backward GEMMs (dX = dZ*W, dW = dZ^T*H) read the same row-major activations / weights as the forward
// pass, so no transposed copies exist anywhere.  Shared-memory layouts follow the canonical UMMA
// SWIZZLE_128B forms documented in CUTLASS cute/atom/mma_traits_sm100.hpp:
//   K-major : rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO)            -> one TMA box {64, rows}
//   MN-major: atoms of [64 k-rows x 64 mn] (k-row = 128 B, 8-row groups SBO=1024 B apart), atoms of
//             successive 64-wide mn chunks LBO = 8192 B apart                      -> one TMA box {64,64} per atom
#include <stdarg.h>
#include <stdlib.h>

#include "hg_common.cuh"

namespace hg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*alignment slack*/ - 256 /*barriers*/;
#ifndef HG_EPI_WARPS
#define HG_EPI_WARPS 16
#endif
#ifndef HG_F32_EPI_WARPS
#define HG_F32_EPI_WARPS 8
#endif
constexpr int MN_ATOM_BYTES = BLOCK_K * 128;  // 8192: one [64 x 64] bf16 atom
constexpr int HG_MAX_COLSUM_COLS = 2048;      // columns (summed over the problems of a group) a DELU launch can column-sum

template <int BLOCK_N, int EPI, int CG>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;                 // 16 KB
  static constexpr int B_BYTES = (BLOCK_N / CG) * BLOCK_K * 2;          // 8 / 16 / 32 KB (per CTA)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // per-warp staging tile: 32 x 36 floats for the transposing split-K epilogue; for the bf16 ones 32 rows x 64 B, dense, in
  // the SWIZZLE_64B layout (16-byte chunk index XOR (row / 2) % 4: conflict-free 16-byte accesses) -- the source of the
  // TMA store of the chunk.  Behind ALL tiles: 256 B per warp for its slice of the bias vector.
  static constexpr int EPI_TILE_BYTES = (EPI == HG_EPI_ATOMIC_F32 || EPI == HG_EPI_STORE_F32) ? 32 * 36 * 4 : 32 * 64;
  static constexpr int BIAS_SLICE_BYTES = (EPI == HG_EPI_ATOMIC_F32 || EPI == HG_EPI_STORE_F32) ? 0 : 256;
  static constexpr bool F32_OUT = (EPI == HG_EPI_ATOMIC_F32 || EPI == HG_EPI_STORE_F32);
  // epilogue warps: 4 per scheduler for the bf16 epilogues (they bound the short-K launches: 16 beat 8 by ~20 %);
  // the fp32 ones sit behind K = rows main loops and keep 8 so that their 4.5 KB staging tiles leave a deeper TMA ring
  // (16, -DHG_F32_EPI_WARPS=16: dW 1.498 -> 1.476 ms per iteration, within noise -- the exposed part of that epilogue is
  // the L2's fp32 reduction rate, not the warps issuing it)
  static constexpr int EPI_WARPS = F32_OUT ? HG_F32_EPI_WARPS : HG_EPI_WARPS;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int COL_SPLIT = EPI_WARPS / 4;       // column slices per tile
  // HEAD: [16][128] fp32 output weights + [4 slices][128][17] fp32 partial sums + per-output constants
  static constexpr int HEAD_BYTES = (16 * 128 + 4 * 128 * 17 + 64) * 4;
  // DELU: per-CTA column-sum accumulators (bias gradients) for every column of the group, flushed once per launch
  // (one array per TMEM lane quarter: slot (quarter, column) has exactly one owner warp, so no shared atomics -- fp32
  // shared atomics are CAS loops)
  static constexpr int COLSUM_BYTES = (EPI == HG_EPI_DELU_BF16) ? 4 * HG_MAX_COLSUM_COLS * 4 : 0;
  static constexpr int EPI_BYTES = (EPI == HG_EPI_BIAS_ELU_HEAD) ? HEAD_BYTES : EPI_WARPS * (EPI_TILE_BYTES + BIAS_SLICE_BYTES) + COLSUM_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET - 768 - EPI_BYTES) / STAGE_BYTES;  // 3..7 deep TMA ring
  // ring | barriers (256 B, padded to 1 KB so that the staging tiles stay 1024-byte aligned) | tiles | bias slices | column sums
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 + 1024 + EPI_BYTES;
  static_assert(STAGES >= 3, "pipeline too shallow");
};

struct DevProblem {
  int M, N, K;
  int a_mn, b_mn;
  int k_splits, kb_per_split, num_kb;
  int tiles_m, tiles_n;
  unsigned inv_tiles_m, inv_tiles_n;   // ceil(2^32 / d): x / d == __umulhi(x, inv) for the tile counts that occur (x * d < 2^32)
  int tile_begin, tile_end;
  void* out;
  long long ldo;
  const float* bias;
  const __nv_bfloat16* act;
  long long ldact;
  float* colsum;
  int colsum_off;      // first slot of this problem in the CTA's shared column-sum array
  int out_fs, out_fd;  // HG_EPI_ATOMIC_F32: frame remap of the output columns (0 = off), see HgGemmProblem
};

struct alignas(64) GroupParams {
  CUtensorMap tma_a[HG_MAX_GROUP];
  CUtensorMap tma_b[HG_MAX_GROUP];
  CUtensorMap tma_out[HG_MAX_GROUP];   // bf16 epilogues: output [M, N], box {32 columns, 32 rows}, SWIZZLE_64B (the chunk of one warp)
  DevProblem p[HG_MAX_GROUP];
  int num_problems;
  int total_tiles;
  // m_major (experiment switch, off by default): tiles numbered (m block, problem, n block) instead of (problem, m block,
  // n block), possible when every problem has the same number of row blocks and no split-K: CTAs working side by side then
  // sit on the SAME rows of different nets (an input shared by two nets is fetched from HBM once).
  int m_major, tiles_per_m;
  // HG_EPI_BIAS_ELU_HEAD only: the fused output layers and the sampling stream of this launch
  HgFusedHead heads[HG_MAX_GROUP];
  const float* eps;
  unsigned long long seed, counter;
};

struct Tile {
  int g, m_blk, n_blk, kb_begin, kb_end;
};

__device__ __forceinline__ Tile decode_tile(const GroupParams& P, int t) {
  if (P.m_major) {
    Tile ti;
    ti.m_blk = t / P.tiles_per_m;
    int r = t - ti.m_blk * P.tiles_per_m, g = 0;
    while (g + 1 < P.num_problems && r >= P.p[g].tiles_n) { r -= P.p[g].tiles_n; ++g; }
    ti.g = g;
    ti.n_blk = r;
    ti.kb_begin = 0;
    ti.kb_end = P.p[g].num_kb;
    return ti;
  }
  int g = 0;
  while (g + 1 < P.num_problems && t >= P.p[g].tile_end) ++g;
  const DevProblem& q = P.p[g];
  // (every warp of the CTA decodes every tile: the four runtime divisions this replaces were ~250 instructions and
  // 10 % of the epilogue samples of the rollout launches)
  const unsigned local = (unsigned)(t - q.tile_begin);
  Tile ti;
  ti.g = g;
  const unsigned l1 = q.tiles_n == 1 ? local : __umulhi(local, q.inv_tiles_n);
  ti.n_blk = (int)(local - l1 * (unsigned)q.tiles_n);
  const unsigned ks_u = q.tiles_m == 1 ? l1 : __umulhi(l1, q.inv_tiles_m);
  ti.m_blk = (int)(l1 - ks_u * (unsigned)q.tiles_m);
  const int ks = (int)ks_u;
  ti.kb_begin = ks * q.kb_per_split;
  ti.kb_end = min(q.num_kb, ti.kb_begin + q.kb_per_split);
  return ti;
}

// Tile `it` of CTA (pair) `group_id`: waves of num_groups tiles, odd waves dealt in reverse ("snake"), so the CTAs that drew
// the expensive tiles of one wave draw the cheap ones of the next.  Problems are listed in decreasing K (actor before
// critic), e.g. the rollout's first layer as 256 x 256 pair tiles: 32 actor tiles (12 k-blocks) + 48 critic tiles (4) on 74
// pairs -- plain round-robin gives six pairs 12 + 4 k-blocks, the snake gives them 4 + 4 and nobody more than 12.
__device__ __forceinline__ int snake_tile(int it, int group_id, int num_groups) {
  return it * num_groups + ((it & 1) ? num_groups - 1 - group_id : group_id);
}

// N of the MMA instruction for this tile = its valid columns rounded up to 32: a ragged last tile (N=705 -> 193,
// the 128-wide lin-vel layer in a 256-wide tile) does not pay for columns that are all zero padding.
template <int BLOCK_N>
__device__ __forceinline__ int tile_n_mma(const DevProblem& q, const Tile& ti) {
  const int n_left = q.N - ti.n_blk * BLOCK_N;
  return n_left >= BLOCK_N ? BLOCK_N : ((n_left + 31) & ~31);
}

__device__ __forceinline__ uint64_t desc_base() {
  // version=1 (bits 46-47), layout SWIZZLE_128B=2 (bits 61-63), SBO = 1024 B (bits 32-45, >>4)
  return (uint64_t(1) << 46) | (uint64_t(2) << 61) | (uint64_t(1024 >> 4) << 32);
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return desc_base() | (uint64_t(1) << 16) | uint64_t((saddr & 0x3FFFFu) >> 4);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return desc_base() | (uint64_t(MN_ATOM_BYTES >> 4) << 16) | uint64_t((saddr & 0x3FFFFu) >> 4);
}

// ------------------------------------------------------------------------------------------ epilogues
// Column sums over the 32 rows held by a warp (one row per lane, 32 columns in registers).
// Butterfly "transpose-reduce": 31 shuffles; on return lane L holds the sum of column L in d[0].
__device__ __forceinline__ float warp_colsum32(float (&d)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float send = upper ? d[i] : d[i + off];
      float keep = upper ? d[i + off] : d[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return d[0];
}

__device__ __forceinline__ float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 w;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(saddr) : "memory");
  return w;
}

// Byte offset of 16-byte chunk `c16` (0..3) of row `row` (0..31) in a warp's staged 32 x 32 bf16 tile: dense 64-byte rows
// in the TMA SWIZZLE_64B pattern (chunk index XOR bits 7-8 of the address = (row / 2) % 4).  Eight lanes of a 16-byte
// access phase hit eight different 16-byte bank groups whether they walk rows (one row per lane) or a row's chunks.
__device__ __forceinline__ uint32_t epi_off(int row, int c16) { return (uint32_t)(row * 64 + ((c16 ^ ((row >> 1) & 3)) << 4)); }

// TMA store of a staged tile: [32 rows x 32 columns] bf16 at (col, row) of the output tensor; rows / columns outside the
// tensor are clipped by the hardware.  One lane issues it after every lane has fenced its shared-memory writes.
__device__ __forceinline__ void tma_store_tile(const void* tmap, uint32_t src, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(col), "r"(row) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the staged tile may be overwritten once the bulk store has READ it
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tile_colsum_mma(uint32_t tile_addr, int lane, float (&c)[4][2]);

// Write the warp's 32 rows x 32 columns (one row per lane, fp32 in registers) as bf16: every lane parks its 64-byte row
// in the warp's staged tile (SWIZZLE_64B layout), one lane hands the tile to the TMA unit (`cp.async.bulk.tensor` store:
// UTMASTG).  The four-lanes-per-row LDS + STG.128 pass this replaces cost ~90 issue slots per chunk in an issue-bound
// epilogue and eight half-line requests per store instruction in the LSU; the bulk store writes whole rows and clips
// rows >= M / columns >= N itself.  The caller has made sure the previous bulk store has read the tile (tma_store_wait_read).
__device__ __forceinline__ void store_rows_bf16(const float (&z)[32], float* tile, int lane, const void* tmap_out, int col, int row_w,
                                                float* colsum = nullptr) {
  const uint32_t tb = smem_u32(tile);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const uint4 w = make_uint4(pack_bf16x2(z[j], z[j + 1]), pack_bf16x2(z[j + 2], z[j + 3]),
                               pack_bf16x2(z[j + 4], z[j + 5]), pack_bf16x2(z[j + 6], z[j + 7]));
    sts128(tb + epi_off(lane, j >> 3), w);
  }
  fence_proxy_async_smem();            // this lane's generic-proxy writes -> visible to the bulk (async-proxy) store
  __syncwarp();
  if (lane == 0) tma_store_tile(tmap_out, tb, col, row_w);
  if (colsum) {     // (warp-uniform) bias gradient of the layer below: column sums of the staged tile, 32 REDs per chunk
    float c[4][2];
    tile_colsum_mma(tb, lane, c);
    // column t lives in c[t / 8][t % 2] of the lanes with lane % 4 == (t % 8) / 2: bring it to lane t (one conflict-free
    // shared atomic per chunk).  Global REDs straight from here are costly: every CTA hits the same few lines of the
    // bias-gradient vector and the L2 serialises requests per line (eight 4-lane REDs per chunk: 2x slower launch).
    float mine = 0.f;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float x = __shfl_sync(0xffffffffu, c[nb][e], (lane & 7) >> 1);
        if ((lane >> 3) == nb && (lane & 1) == e) mine = x;
      }
    }
    colsum[lane] += mine;                 // this warp's own shared-memory accumulator (flushed once at the end of the launch)
  }
}

// Column sums of the warp's staged 32 x 32 bf16 tile (SWIZZLE_64B layout, see epi_off) on the tensor core: ones[16x16] x
// tile[16 x 8] per mma.sync.m16n8k16, the tile read as B fragments with ldmatrix.trans.  Every row of the product holds
// the column sums, so lane t ends up with the sums of columns nb*8 + 2*(t%4) + {0,1}, nb = 0..3, in c[nb][0..1].
// Replaces a 31-shuffle butterfly (~190 instructions per chunk) with 4 ldmatrix + 8 mma + 8 shuffles; the summands are
// the bf16-rounded values the weight-gradient GEMM reads, accumulated in fp32.
__device__ __forceinline__ void tile_colsum_mma(uint32_t tile_addr, int lane, float (&c)[4][2]) {
  const uint32_t ones = 0x3F803F80u;      // bf16x2 (1.0, 1.0)
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    uint32_t b0, b1, b2, b3;              // k rows 0-7, 8-15, 16-23, 24-31 of column block nb
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(tile_addr + epi_off(lane, nb)) : "memory");
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};"
                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(ones), "r"(b0), "r"(b1));
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %4, %4, %4}, {%5, %6}, {%0, %1, %2, %3};"
                 : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3) : "r"(ones), "r"(b2), "r"(b3));
    c[nb][0] = d0;
    c[nb][1] = d1;
  }
}

// Activation rows of the layer below for one 32x32 chunk, in the coalesced pattern of store_rows_bf16 (4 lanes per
// 64-byte row segment).  Issued BEFORE the accumulator barrier so the HBM latency hides behind the tile's main loop.
// `act_p`: this lane's first 16-byte piece of the chunk (see store_rows_bf16); `live`: the chunk is a full chunk of the problem.
__device__ __forceinline__ void prefetch_act(const uint4* act_p, long long ld_u4, int rows_left, bool live, int lane, uint4 (&pre)[4]) {
  const int r0 = lane >> 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pre[i] = make_uint4(0, 0, 0, 0);
    if (live && r0 + 8 * i < rows_left) pre[i] = __ldg(act_p + i * ld_u4);
  }
}

// Per-(warp, tile) addressing shared by the chunks of a tile, computed once after the tile decode.
struct EpiCtx {
  const void* tmap_out;  // bf16 outputs: the problem's output tensor map (TMA store of the warp's chunks)
  int row_w;             // first of the warp's 32 rows
  const uint4* act_p;    // DELU: this lane's first 16-byte piece of the activation tile (chunk c: + 4 c)
  long long ldact_u4;    // 8 rows further, in 16-byte units
  int rows_left;         // valid rows of the warp's 32 (<= 0: none)
};

template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const DevProblem& q, int row, int n0, uint32_t (&v)[32], int lane, float* tile,
                                               const uint4 (&pre)[4], float* s_colsum, const float* bias_s, const EpiCtx& X, int c) {
  const bool row_ok = row < q.M;
  const bool full = (n0 + 32 <= q.N);
  if constexpr (EPI == HG_EPI_BIAS_ELU_BF16 || EPI == HG_EPI_BIAS_BF16 || EPI == HG_EPI_DELU_BF16) {
    if (lane == 0) tma_store_wait_read();      // the previous chunk's bulk store has read the staged tile (long done by now)
    __syncwarp();
  }
  if constexpr (EPI == HG_EPI_STORE_F32) {
    float* o = reinterpret_cast<float*>(q.out) + (long long)row * q.ldo + n0;
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < q.N) o[j] = __uint_as_float(v[j]) + (q.bias ? __ldg(q.bias + n0 + j) : 0.f);
    }
  } else if constexpr (EPI == HG_EPI_ATOMIC_F32) {
    // Transpose the warp's 32x32 chunk through shared memory (rows 36 floats apart: 16-byte aligned and conflict-free
    // both ways) so that every RED instruction adds CONSECUTIVE floats of an output row.  Whole chunks of a 16-byte
    // aligned output go out as red.global.add.v4.f32: 8 lanes cover a row's 128 bytes, 8 instructions the chunk
    // (a quarter of the scalar form's LSU work); ragged chunks and the frame-remapped layer-1 gradients take the
    // scalar form, 32 consecutive floats per instruction.
    const uint32_t tile_s = smem_u32(tile);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts128(tile_s + lane * 144 + j * 16, make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    __syncwarp();
    const int row_base = row - lane;
    float* out = reinterpret_cast<float*>(q.out);
    if (full && !q.out_fs && ((q.ldo & 3) | (int)(reinterpret_cast<uintptr_t>(out) & 15)) == 0) {
      const int r0 = lane >> 3, c4 = lane & 7;
      float* o = out + (long long)(row_base + r0) * q.ldo + n0 + c4 * 4;
      const long long step = 4ll * q.ldo;
#pragma unroll
      for (int i = 0; i < 8; ++i, o += step) {
        const uint4 w = lds128(tile_s + (i * 4 + r0) * 144 + c4 * 16);
        if (row_base + i * 4 + r0 < q.M)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(__uint_as_float(w.x)), "f"(__uint_as_float(w.y)),
                       "f"(__uint_as_float(w.z)), "f"(__uint_as_float(w.w)) : "memory");
      }
    } else {
      bool col_ok = n0 + lane < q.N;
      int ocol = n0 + lane;
      if (q.out_fs) {      // frame-pitched input columns -> the packed columns of the weight gradient; pad columns are dropped
        const int f = ocol / q.out_fs, c = ocol - f * q.out_fs;
        col_ok = col_ok && c < q.out_fd;
        ocol = f * q.out_fd + c;
      }
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        if (col_ok && row_base + i < q.M) atomicAdd(out + (long long)(row_base + i) * q.ldo + ocol, tile[i * 36 + lane]);
      }
    }
    __syncwarp();
  } else if constexpr (EPI == HG_EPI_BIAS_ELU_BF16 || EPI == HG_EPI_BIAS_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(q.out) + (long long)row * q.ldo + n0;
    float z[32], bv[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {        // the chunk's biases out of the warp's shared-memory copy (zero past N / without bias)
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
      bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = __uint_as_float(v[j]) + bv[j];
#ifdef HG_DBG_NOEXP
      z[j] = t;
#else
      // ELU with MUFU.EX2: |err| ~ 1e-7 absolute, far below the bf16 rounding of the stored activation
      z[j] = (EPI == HG_EPI_BIAS_ELU_BF16) ? (t > 0.f ? t : exp2f_fast(t * 1.4426950408889634f) - 1.f) : t;
#endif
    }
    if (full) {
      store_rows_bf16(z, tile, lane, X.tmap_out, n0, X.row_w);
    } else if (row_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < q.N) o[j] = __float2bfloat16_rn(z[j]);
    }
  } else if constexpr (EPI == HG_EPI_DELU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(q.out) + (long long)row * q.ldo + n0;
    const __nv_bfloat16* a = q.act + (long long)row * q.ldact + n0;
    float d[32];
    if (full) {
      // activation rows of the layer below: coalesced global -> shared (4 lanes per 64-byte row segment), then each
      // lane reads back its own row
      const uint32_t tb = smem_u32(tile);
#ifndef HG_DX_DIRECT_ACT
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), piece = lane & 3;
        sts128(tb + epi_off(r, piece), pre[i]);
      }
      __syncwarp();
#endif
      // dZ = dH * elu'(h), elu'(h) = 1 (h > 0) or h + 1 (h <= 0)  ==  fma(dH, min(h, 0), dH).  Rows past M need no
      // predicate: TMA zero-fills their A rows, so their accumulators are exactly 0.
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#ifdef HG_DX_DIRECT_ACT
        const uint4 hw = pre[j >> 3];
#else
        const uint4 hw = lds128(tb + epi_off(lane, j >> 3));
#endif
        const uint32_t hv[4] = {hw.x, hw.y, hw.z, hw.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t m2;                                   // min(h, 0) on the packed pair, then widen (bf16 -> fp32 is a shift)
          asm("min.bf16x2 %0, %1, %2;" : "=r"(m2) : "r"(hv[k]), "r"(0u));
          const float a0 = __uint_as_float(v[j + 2 * k]), a1 = __uint_as_float(v[j + 2 * k + 1]);
          d[j + 2 * k] = fmaf(a0, __uint_as_float(m2 << 16), a0);
          d[j + 2 * k + 1] = fmaf(a1, __uint_as_float(m2 & 0xFFFF0000u), a1);
        }
      }
      __syncwarp();
#ifndef HG_COLSUM_BUTTERFLY
      store_rows_bf16(d, tile, lane, X.tmap_out, n0, X.row_w, q.colsum ? s_colsum + q.colsum_off + n0 : nullptr);
      return;
#else     // A/B switch (tools/build_variant.sh): fp32 shuffle butterfly below
      store_rows_bf16(d, tile, lane, X.tmap_out, n0, X.row_w);
#endif
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float val = 0.f;
        if (row_ok && n0 + j < q.N) {
          float h = __bfloat162float(a[j]);
          val = __uint_as_float(v[j]) * (h > 0.f ? 1.f : h + 1.f);
          o[j] = __float2bfloat16_rn(val);
        }
        d[j] = val;
      }
    }
    if (q.colsum) {
      // bias gradient of the layer below = column sums of what was just written (bf16-rounded values
      // are what the weight-gradient GEMM will read, but the fp32 sum is the better estimate).
      float s = warp_colsum32(d, lane);
      if (n0 + lane < q.N) s_colsum[q.colsum_off + n0 + lane] += s;
    }
  }
}


// ------------------------------------------------------------------------------------------ fused rollout head
// Epilogue of HG_EPI_BIAS_ELU_HEAD (16 epilogue warps = 4 TMEM lane quarters x 4 column slices of 32): every warp turns
// its 32 x 32 accumulator chunk into h = bf16(elu(acc + bias)), multiplies it with the staged fp32 output weights and
// writes its partial sums to the tile's [slice][128 rows][n_out] array in shared memory; after a barrier among the
// epilogue warps, one thread per row adds the four slices IN A FIXED ORDER (rollouts stay bit-reproducible) and
// finishes the head (policy: sample + log-prob; linear: value).
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <int BLOCK_N, int CG>
__device__ __forceinline__ void head_epilogue(const GroupParams& P, float* sm, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                              uint32_t tmem_base, int group_id, int num_groups, int warp, int lane) {
  static_assert(BLOCK_N == 128 || CG >= 1, "");
  float* sW = sm;                    // [16][128]
  float* sAcc = sm + 16 * 128;       // [4][128][17]
  float* sB = sAcc + 4 * 128 * 17;   // [16] output bias, [16] std
  const int quad = warp & 3, slice = (warp - 4) >> 2;
  const int et = (warp - 4) * 32 + lane;       // 0..511
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int it = 0, t; (t = snake_tile(it, group_id, num_groups)) < P.total_tiles; ++it) {
    const Tile ti = decode_tile(P, t);
    const DevProblem& q = P.p[ti.g];
    const HgFusedHead& hd = P.heads[ti.g];
    const int n_out = hd.n_out;
    for (int i = et; i < n_out * q.N; i += 512) {
      const int a = i / q.N, k = i - a * q.N;
      sW[a * 128 + k] = __ldg(hd.W + i);
    }
    if (et < n_out) {
      sB[et] = __ldg(hd.b + et);
      sB[16 + et] = hd.kind == HG_HEAD_POLICY ? __ldg(hd.std + et) : 0.f;
    }
    epi_bar_sync();
    mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
    tc_fence_after();
    const int n0 = slice * 32;
    const int rt = quad * 32 + lane;           // row inside the tile
    const bool live = n0 < q.N;                // N % 32 == 0: a live chunk is a full chunk
    uint32_t v[32];
    if (live) {
      tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + n0, v);
      tmem_ld_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));      // accumulator is in registers: hand it back
    if (live) {
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(v[j]) + (q.bias ? __ldg(q.bias + n0 + j) : 0.f);
        const float e = x > 0.f ? x : exp2f_fast(x * 1.4426950408889634f) - 1.f;
        z[j] = __bfloat162float(__float2bfloat16_rn(e));     // what the stored activation would hold
      }
      for (int a = 0; a < n_out; ++a) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + a * 128 + n0);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 wa = w4[j >> 2], wb = w4[(j >> 2) + 1];
          s0 += z[j] * wa.x + z[j + 1] * wa.y + z[j + 2] * wa.z + z[j + 3] * wa.w;
          s1 += z[j + 4] * wb.x + z[j + 5] * wb.y + z[j + 6] * wb.z + z[j + 7] * wb.w;
        }
        sAcc[(slice * 128 + rt) * 17 + a] = s0 + s1;
      }
    } else {
      for (int a = 0; a < n_out; ++a) sAcc[(slice * 128 + rt) * 17 + a] = 0.f;
    }
    epi_bar_sync();
    const int row = ti.m_blk * BLOCK_M + rt;
    // every thread of the row's four slices finishes `per` outputs (fixed-order sum of the slice partials), so the
    // sampling / log-prob math is spread over all 512 epilogue threads
    auto dot = [&](int a) {
      return ((sAcc[rt * 17 + a] + sAcc[(128 + rt) * 17 + a]) + sAcc[(256 + rt) * 17 + a]) + sAcc[(384 + rt) * 17 + a];
    };
    const int per = (n_out + 3) >> 2, a_lo = slice * per, a_hi = min(n_out, a_lo + per);
    float lp = 0.f;
    if (row < q.M) {
      if (hd.kind == HG_HEAD_POLICY) {
        int have = -1;
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        for (int a = a_lo; a < a_hi; ++a) {
          if (P.eps == nullptr && (a >> 2) != have) {      // four normals per Philox call, keyed by (row, a / 4)
            have = a >> 2;
            const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)have, (uint32_t)P.counter, (uint32_t)(P.counter >> 32)),
                                          make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
            const float2 n01 = box_muller(r.x, r.y), n23 = box_muller(r.z, r.w);
            e[0] = n01.x; e[1] = n01.y; e[2] = n23.x; e[3] = n23.y;
          }
          const int k = a & 3;
          const float m = dot(a) + sB[a], sd = sB[16 + a];
          const float ee = P.eps ? __ldg(P.eps + (long long)row * n_out + a) : (k == 0 ? e[0] : k == 1 ? e[1] : k == 2 ? e[2] : e[3]);
          const float act = m + sd * ee;                         // Normal(mu, std).sample()  (actor_critic.py:129,133)
          const float d = act - m;
          lp += -(d * d) / (2.f * sd * sd) - logf(sd) - LOG_SQRT_2PI;   // Normal.log_prob
          hd.actions[(long long)row * n_out + a] = act;
          hd.out[(long long)row * n_out + a] = m;
          hd.sigma[(long long)row * n_out + a] = sd;
        }
      } else {
        for (int a = a_lo; a < a_hi; ++a) hd.out[(long long)row * n_out + a] = dot(a) + sB[a];
      }
    }
    if (hd.kind == HG_HEAD_POLICY) {       // (uniform per tile) log-prob = fixed-order sum of the four slices' terms
      epi_bar_sync();                      // every dot() has been read: the partial-sum array can be reused
      sAcc[slice * 128 + rt] = lp;
      epi_bar_sync();
      if (slice == 0 && row < q.M) hd.log_prob[row] = ((sAcc[rt] + sAcc[128 + rt]) + sAcc[256 + rt]) + sAcc[384 + rt];
    }
    epi_bar_sync();          // the next tile re-stages sW / sAcc
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

// ------------------------------------------------------------------------------------------ cta_group helpers
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-pair rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  if constexpr (CG == 1) {
    tma_load_2d(dst, tmap, bar, c0, c1);
  } else {
    // executed by both CTAs of the pair; the transaction bytes land on the LEADER's barrier (peer bit cleared)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive (once all earlier MMAs of this thread retired) on the barrier at the same offset in EVERY CTA of the group
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint32_t bar) {
  if constexpr (CG == 1) {
    umma_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)0x3)
                 : "memory");
  }
}
template <int CG, int kCols>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t smem_dst) {
  if constexpr (CG == 1) {
    tmem_alloc<kCols>(smem_dst);
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG, int kCols>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr) {
  if constexpr (CG == 1) {
    tmem_dealloc<kCols>(taddr);
  } else {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ kernel
// CG = 1: one CTA per 128 x BLOCK_N tile.
// CG = 2: a CTA pair (cluster of 2, `tcgen05.mma.cta_group::2`) per 256 x BLOCK_N tile: CTA r owns rows
//         [128 r, 128 r + 128) of A and of the accumulator and loads only columns [r BLOCK_N/2, (r+1) BLOCK_N/2) of B;
//         the leader CTA issues the MMAs for both, each tensor core reads the other half of B from its peer's
//         shared memory.  Operand fill per SM drops from 16 + 32 KB to 16 + 16 KB per 128x256x64 MMA block.
// STEP = true: the body runs as the layer-1 part of rollout_step_kernel (below): CTA `group_id` of `num_groups`, and every
// epilogue warp publishes each finished tile in `step_flags[problem * tiles_m + m_blk]` once its bulk stores have completed.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int BLOCK_N, int EPI, int CG, bool STEP>
__device__ __forceinline__ void gemm_body(const GroupParams& P, const int group_id, const int num_groups, unsigned* step_flags) {
  using K = Cfg<BLOCK_N, EPI, CG>;
  constexpr int A_BYTES = K::A_BYTES, STAGE_BYTES = K::STAGE_BYTES, STAGES = K::STAGES;
  constexpr int B_ROWS = BLOCK_N / CG;                      // rows (K-major) / columns (MN-major) of B this CTA loads
  constexpr int TMEM_COLS = 2 * BLOCK_N;                    // 256 / 512
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]   (CG=2: only the leader's are used)
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]  (CG=2: only the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* epi_tiles = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 1024);     // 1024-byte aligned (TMA store source)

  pdl_trigger();
  if (threadIdx.x == 0) trace(0x200);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < P.num_problems; ++g) {
      tma_prefetch_desc(&P.tma_a[g]);
      tma_prefetch_desc(&P.tma_b[g]);
      if constexpr (EPI == HG_EPI_BIAS_ELU_BF16 || EPI == HG_EPI_BIAS_BF16 || EPI == HG_EPI_DELU_BF16) tma_prefetch_desc(&P.tma_out[g]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), CG * K::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg<CG, TMEM_COLS>(smem_u32(tmem_slot));
  if (warp == 3 && lane < HG_MAX_GROUP) {
    // (idle warp) pull the problem table through the constant cache while the predecessor still runs: the device timeline
    // showed 1.1 us between the end of the set-up and the first TMA issue -- the producer's tile decode missing on
    // every dependent parameter read
    const DevProblem& w = P.p[lane];
    const int touch = w.M + w.tile_end + w.out_fd + (int)(reinterpret_cast<uintptr_t>(w.out) ^ reinterpret_cast<uintptr_t>(w.colsum)) +
                      P.num_problems + P.total_tiles + P.tiles_per_m + (int)P.counter + P.heads[lane].n_out;
    asm volatile("" ::"r"(touch));
  }
  tc_fence_before();
  pdl_wait();          // everything above is on-chip set-up; operands, bias, act and outputs are touched only below
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace(0x201);

  if (warp == 0) {
    // ================================================================ TMA producer (every CTA)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0, t; (t = snake_tile(it, group_id, num_groups)) < P.total_tiles; ++it) {
        const Tile ti = decode_tile(P, t);
        const DevProblem& q = P.p[ti.g];
        const void* ta = &P.tma_a[ti.g];
        const void* tb = &P.tma_b[ti.g];
        const int m0 = (ti.m_blk * CG + rank) * BLOCK_M;
        // CTA `rank` of a pair supplies columns [rank * n_mma/2, (rank+1) * n_mma/2) of the tile's B operand
        const int nb0 = ti.n_blk * BLOCK_N + rank * (tile_n_mma<BLOCK_N>(q, ti) / CG);
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if (rank == 0) mbar_arrive_expect_tx(fb, CG * STAGE_BYTES);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (!q.a_mn) {
            tma_load_2d_cg<CG>(sa, ta, fb, kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d_cg<CG>(sa + j * MN_ATOM_BYTES, ta, fb, m0 + j * 64, kb * BLOCK_K);
          }
          if (!q.b_mn) {
            tma_load_2d_cg<CG>(sb, tb, fb, kb * BLOCK_K, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < B_ROWS / 64; ++j)
              tma_load_2d_cg<CG>(sb + j * MN_ATOM_BYTES, tb, fb, nb0 + j * 64, kb * BLOCK_K);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (it == 0 && kb == ti.kb_begin) trace(0x202);
        }
      }
      // every operand load of this CTA is in flight: what remains is its last MMAs and epilogue.  The dependent launch is
      // triggered once every CTA of the (fully resident, persistent) grid is here -- its launch latency hides behind that tail.
      pdl_launch_dependents();
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0, t; (t = snake_tile(it, group_id, num_groups)) < P.total_tiles; ++it) {
        const Tile ti = decode_tile(P, t);
        const DevProblem& q = P.p[ti.g];
        // instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), majors, N>>3 at 17, M>>4 at 24
        const int n_mma = tile_n_mma<BLOCK_N>(q, ti);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(q.a_mn) << 15) |
                               (uint32_t(q.b_mn) << 16) | (uint32_t(n_mma >> 3) << 17) |
                               (uint32_t((BLOCK_M * CG) >> 4) << 24);
        // Descriptors are built with 32-bit arithmetic only: the high word (SBO, version, swizzle) is a constant and the
        // low word is (LBO << 16) | (address >> 4), advanced by a constant per 16-wide K slice -- the single issuing
        // thread must stay far ahead of the tensor core (a 128x128x16 MMA retires in 64 clocks).
        const uint32_t desc_hi = (uint32_t)(desc_base() >> 32);
        const uint32_t a_lo0 = q.a_mn ? (uint32_t(MN_ATOM_BYTES >> 4) << 16) : (1u << 16);
        const uint32_t b_lo0 = q.b_mn ? (uint32_t(MN_ATOM_BYTES >> 4) << 16) : (1u << 16);
        const uint32_t a_step = q.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;   // MN-major: 16 k-rows; K-major: 32 B
        const uint32_t b_step = q.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        const int K_total = q.K;
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        for (int kb = ti.kb_begin; kb < ti.kb_end; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          if (it == 0 && kb == ti.kb_begin) trace(0x203);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_lo = a_lo0 | ((sa & 0x3FFFFu) >> 4);
          const uint32_t b_lo = b_lo0 | (((sa + A_BYTES) & 0x3FFFFu) >> 4);
          const int k_left = K_total - kb * BLOCK_K;
          if (k_left >= BLOCK_K) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t ad = (uint64_t(desc_hi) << 32) | (a_lo + k * a_step);
              const uint64_t bd = (uint64_t(desc_hi) << 32) | (b_lo + k * b_step);
              umma_bf16_cg<CG>(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
            }
          } else {
            const int nk = (k_left + UMMA_K - 1) / UMMA_K;
            for (int k = 0; k < nk; ++k) {
              const uint64_t ad = (uint64_t(desc_hi) << 32) | (a_lo + k * a_step);
              const uint64_t bd = (uint64_t(desc_hi) << 32) | (b_lo + k * b_step);
              umma_bf16_cg<CG>(d_tmem, ad, bd, idesc, k == 0 ? accumulate : 1u);
            }
          }
          accumulate = 1;
          umma_commit_cg<CG>(smem_u32(&empty_bar[stage]));   // smem slot free (in every CTA of the group) once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg<CG>(smem_u32(&tfull_bar[acc]));        // accumulator complete -> epilogues
        if (it < 4) trace(0x204, it);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue: warps 4..(4+NUM_EPI_WARPS)
    const int quad = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;          // which column slice of the tile
    constexpr int SLICE = BLOCK_N / K::COL_SPLIT;
    constexpr int CHUNKS = SLICE / 32;         // 32-column chunks per warp
    float* tile = epi_tiles + (warp - 4) * (K::EPI_TILE_BYTES / 4);
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI == HG_EPI_BIAS_ELU_HEAD) {
      head_epilogue<BLOCK_N, CG>(P, epi_tiles, tfull_bar, tempty_bar, tmem_base, group_id, num_groups, warp, lane);
    } else {
    // Measured on B200 (tools/bench_gemm_dx.py, 61440 x {512,768} x 256): sixteen epilogue warps with the plain loop
    // below beat eight by ~20 %; double-buffering the TMEM loads or fetching the DELU activations a tile ahead costs
    // more in registers (640 threads -> 102 per thread) than it hides in latency, so neither is done.
    constexpr bool DELU = (EPI == HG_EPI_DELU_BF16);
    float* bias_s = epi_tiles + K::EPI_WARPS * (K::EPI_TILE_BYTES / 4) + (warp - 4) * (K::BIAS_SLICE_BYTES / 4);   // this warp's bias slice
    float* s_colsum_all = epi_tiles + K::EPI_WARPS * ((K::EPI_TILE_BYTES + K::BIAS_SLICE_BYTES) / 4);   // [4 quarters][HG_MAX_COLSUM_COLS]
    float* s_colsum = s_colsum_all + quad * HG_MAX_COLSUM_COLS;   // column c of a problem is always handled by the same
    const int et = (warp - 4) * 32 + lane;                        // (quarter, column slice) warp: plain read-modify-write
    if constexpr (DELU) {
      for (int i = et; i < 4 * HG_MAX_COLSUM_COLS; i += K::EPI_WARPS * 32) s_colsum_all[i] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(K::EPI_WARPS * 32) : "memory");
    }
    for (int it = 0, t; (t = snake_tile(it, group_id, num_groups)) < P.total_tiles; ++it) {
      const Tile ti = decode_tile(P, t);
      const DevProblem& q = P.p[ti.g];
      const int row = (ti.m_blk * CG + rank) * BLOCK_M + quad * 32 + lane;
      const int col0 = ti.n_blk * BLOCK_N + half * SLICE;
      const int nchunks = q.N > col0 ? min(CHUNKS, (q.N - col0 + 31) >> 5) : 0;
      // addressing of this warp's 32 rows x SLICE columns, once per tile (bf16 epilogues; rows are 16-byte aligned: checked
      // by the host)
      EpiCtx X;
      {
        const int row_w = row - lane, r0 = lane >> 2, piece = lane & 3;
        X.row_w = row_w;
        X.rows_left = q.M - row_w;
        X.tmap_out = &P.tma_out[ti.g];
        X.ldact_u4 = q.ldact;          // 8 rows further, in 16-byte units: 8 * ld elements * 2 B / 16 B = ld
        X.act_p = nullptr;
        if constexpr (DELU)
          X.act_p = reinterpret_cast<const uint4*>(q.act + (long long)(row_w + r0) * q.ldact + col0) + piece;
      }
      // DELU: the activation rows are requested BEFORE the accumulator barrier, their latency hides behind the main loop
      uint4 pre[DELU ? CHUNKS : 1][4];
      if constexpr (DELU) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) prefetch_act(X.act_p + 4 * c, X.ldact_u4, X.rows_left, col0 + c * 32 + 32 <= q.N, lane, pre[c]);
      }
      if constexpr (EPI == HG_EPI_BIAS_ELU_BF16 || EPI == HG_EPI_BIAS_BF16) {
        // the slice's biases -> shared memory BEFORE the accumulator is waited for: a global load per chunk sat on the
        // epilogue's critical path (~0.5-1 us per tile; it matters for the rollout launches, whose last epilogue is exposed)
        for (int j = lane; j < SLICE; j += 32) bias_s[j] = (q.bias && col0 + j < q.N) ? __ldg(q.bias + col0 + j) : 0.f;
        __syncwarp();
      }
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      if (warp == 4 && lane == 0 && it < 4) trace(0x205, it);
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + half * SLICE;
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        if (c < nchunks) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (warp == 4 && lane == 0 && it < 2 && c < 2) trace(0x208, it * 2 + c);
#if defined(HG_DBG_EPI_NONE)
#pragma unroll
          for (int j = 0; j < 32; ++j) asm volatile("" ::"r"(v[j]));
#else
          epilogue_chunk<EPI>(q, row, col0 + c * 32, v, lane, tile, pre[DELU ? c : 0], s_colsum, bias_s + c * 32, X, c);
#endif
          if (warp == 4 && lane == 0 && it < 2 && c < 2) trace(0x209, it * 2 + c);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) {
          mbar_arrive(smem_u32(&tempty_bar[acc]));
        } else {   // the leader's MMA thread waits for the epilogues of BOTH CTAs
          asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];"
                       ::"r"(smem_u32(&tempty_bar[acc]) & PEER_BIT_MASK) : "memory");
        }
      }
      if (warp == 4 && lane == 0 && it < 4) trace(0x206, it);
      if constexpr (STEP) {
        // this warp's part of tile (problem, m block) is in global memory: completed bulk stores -> proxy fence ->
        // gpu-scope release of the tile's counter (the tail CTA of that row block acquires it before its TMA loads of h1)
        if (lane == 0) {
          tma_store_wait_all();
          fence_proxy_async_all();
          __threadfence();
          atomicAdd(step_flags + ti.g * q.tiles_m + ti.m_blk, 1u);
        }
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (!K::F32_OUT) {
      if (lane == 0) tma_store_wait_all();     // this warp's bulk stores are complete before the CTA's shared memory goes away
      __syncwarp();
    }
    if constexpr (DELU) {      // one flush per CTA and launch: (columns / 32) coalesced REDs instead of one per 32 x 32 chunk
      asm volatile("bar.sync 1, %0;" ::"n"(K::EPI_WARPS * 32) : "memory");
      for (int g = 0; g < P.num_problems; ++g) {
        const DevProblem& q = P.p[g];
        if (!q.colsum) continue;
        for (int c = et; c < q.N; c += K::EPI_WARPS * 32) {
          const float* sc = s_colsum_all + q.colsum_off + c;
          const float v = ((sc[0] + sc[HG_MAX_COLSUM_COLS]) + sc[2 * HG_MAX_COLSUM_COLS]) + sc[3 * HG_MAX_COLSUM_COLS];
          if (v != 0.f) atomicAdd(q.colsum + c, v);
        }
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(0x207);
  if constexpr (CG == 2) cluster_sync_all();   // nobody leaves while the peer may still read its shared memory / barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg<CG, TMEM_COLS>(tmem_base);
  }
}

template <int BLOCK_N, int EPI, int CG>
__global__ void __launch_bounds__((Cfg<BLOCK_N, EPI, CG>::THREADS), 1) gemm_kernel(const __grid_constant__ GroupParams P) {
  gemm_body<BLOCK_N, EPI, CG, false>(P, blockIdx.x / CG, gridDim.x / CG, nullptr);
}


// ------------------------------------------------------------------------------------------ rollout tail
// Layer 2, layer 3 and the output layer of a net in ONE launch (rollout, three hidden layers): a CTA owns a 128-row
// tile of one net, computes h2 = elu(h1 W2^T + b2) (128 x N2 <= 256, accumulator in TMEM columns 0..255), parks it as
// bf16 in shared memory in the canonical K-major SWIZZLE_128B layout (exactly what a TMA load of an activation tile would
// have produced), multiplies it by W3 (accumulator in TMEM columns 256..383), parks h3 = elu(. + b3) the same way and
// runs the OUTPUT layer on the tensor core as well (W_out split into three bf16 terms, 128 x 16 accumulator in TMEM
// columns 384..399): the device timeline (tools/trace_rollout.py) showed the 12 x 128 fp32 dot products per row on the
// CUDA cores taking 5 us of the actor tile's 9 us head -- the critical path of the launch.  Each thread then finishes
// four outputs of its row (sampling, log-prob).  h2 / h3 never travel through L2 and the step has one launch less.
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-19: the two epilogues
namespace tail {
#ifndef HG_TAIL_STAGES
#define HG_TAIL_STAGES 3
#endif
constexpr int STAGES = HG_TAIL_STAGES;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;            // 16 KB: h1 tile 128 x 64
constexpr int B_BYTES = 256 * BLOCK_K * 2;                // 32 KB: W2 tile 256 x 64
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int KB_BYTES = BLOCK_M * BLOCK_K * 2;           // 16 KB: one 64-wide k-block of h2 (128 rows) / of W3 (<= 128 rows)
constexpr int H2_OFF = STAGES * STAGE_BYTES;              // 144 KB
// h2: 64 KB.  Once layer 3 has read it the region is reused: its first 32 KB hold the h3 tile (bf16, the A operand of the
// output layer), the log-prob partial sums [4 slices][128 rows] sit behind that.
constexpr int LP_OFF = H2_OFF + 2 * KB_BYTES;
// output-layer weights as the B operand of the tensor core: W = hi + mid + lo, three bf16 terms (24 mantissa bits: exact),
// each [16 rows x 128 k] K-major SWIZZLE_128B = 2 k-blocks x 2 KB
constexpr int WOUT_OFF = H2_OFF + 4 * KB_BYTES;
constexpr int WOUT_KB_BYTES = 16 * BLOCK_K * 2;           // 2 KB
constexpr int WOUT_TERM_BYTES = 2 * WOUT_KB_BYTES;        // 4 KB
constexpr int BAR_OFF = WOUT_OFF + 3 * WOUT_TERM_BYTES;
constexpr int HEAD_OFF = BAR_OFF + 256;
constexpr int HEAD_BYTES = (128 + 64 + 256) * 4;          // layer-3 bias [128] + output bias [16] / std [16] / log std [16] / 2 std^2 [16] + layer-2 bias [256]
constexpr int SMEM_TOTAL = HEAD_OFF + HEAD_BYTES + 1024;
constexpr int TMEM_OUT_COL = 384;                         // output-layer accumulator: 16 columns behind layer 2 (0-255) and layer 3 (256-383)
constexpr int THREADS = 128 + 32 * 16;
constexpr int TMEM_COLS = 512;
static_assert(SMEM_TOTAL <= 227 * 1024, "rollout tail: shared memory");
static_assert(LP_OFF + (512 + 3 * 2048) * 4 <= WOUT_OFF, "log-prob sums + staged outputs fit behind the h3 tile");
}  // namespace tail

struct TailNet {
  CUtensorMap tma_h1, tma_w2, tma_w3;
  int M, K1, N2, N3;
  const float* b2;
  const float* b3;
  HgFusedHead head;
};
struct alignas(64) TailParams {
  TailNet net[2];
  int num_nets, tiles_m;
  const float* eps;
  unsigned long long seed, counter;
};

// step_flags != nullptr: the body runs inside rollout_step_kernel and waits for the layer-1 tiles of its row block (counter
// of (net, m block) reaching `step_target`) before it loads h1.
__device__ __forceinline__ void tail_body(const TailParams& P, const int g, const int m_blk, const unsigned* step_flags, const unsigned step_target) {
  using namespace tail;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_bar = bars;            // [STAGES]
  uint64_t* empty_bar = bars + STAGES;  // [STAGES]
  uint64_t* acc1_full = bars + 2 * STAGES;       // layer-2 accumulator complete (MMA -> epilogue, producer)
  uint64_t* h2_ready = bars + 2 * STAGES + 1;    // h2 tile written to shared memory (16 epilogue warps -> MMA)
  uint64_t* w3_full = bars + 2 * STAGES + 2;     // W3 landed in the (idle) operand ring
  uint64_t* acc2_full = bars + 2 * STAGES + 3;   // layer-3 accumulator complete
  uint64_t* h3_ready = bars + 2 * STAGES + 4;    // h3 tile written to shared memory (16 epilogue warps -> MMA)
  uint64_t* acc3_full = bars + 2 * STAGES + 5;   // output-layer accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);
  float* sm = reinterpret_cast<float*>(smem + HEAD_OFF);

  pdl_trigger();
  if (threadIdx.x == 0) trace(0x300);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TailNet& q = P.net[g];
  const int num_kb1 = (q.K1 + BLOCK_K - 1) / BLOCK_K, num_kb2 = q.N2 / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&q.tma_h1);
    tma_prefetch_desc(&q.tma_w2);
    tma_prefetch_desc(&q.tma_w3);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(acc1_full), 1);
    mbar_init(smem_u32(h2_ready), 16);
    mbar_init(smem_u32(w3_full), 1);
    mbar_init(smem_u32(acc2_full), 1);
    mbar_init(smem_u32(h3_ready), 16);
    mbar_init(smem_u32(acc3_full), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  if (warp == 3 && lane == 0) {      // (idle warp) parameters through the constant cache before they are needed, see gemm_kernel
    const int touch = q.M + q.K1 + q.N2 + q.N3 + q.head.n_out + q.head.kind + (int)P.counter +
                      (int)(reinterpret_cast<uintptr_t>(q.b2) ^ reinterpret_cast<uintptr_t>(q.b3) ^ reinterpret_cast<uintptr_t>(q.head.log_prob));
    asm volatile("" ::"r"(touch));
  }
  tc_fence_before();
  pdl_wait();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t ring = smem_u32(smem), h2s = smem_u32(smem + H2_OFF);
  if (threadIdx.x == 0) trace(0x301);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (step_flags) {
        // fused rollout step: h1 of this row block is being produced by the layer-1 CTAs of the SAME launch (they never wait
        // for anything of ours and have lower block indices, i.e. are resident first).  Acquire the tile counter, then order
        // the async-proxy loads behind it.  Watchdog as in mbar_wait: a protocol bug traps instead of hanging.
        const unsigned* f = step_flags + g * P.tiles_m + m_blk;
        const long long t0 = clock64();
        for (;;) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
          if ((int)(v - step_target) >= 0) break;
          if (clock64() - t0 > 4000000000LL) __trap();
          __nanosleep(64);
        }
        fence_proxy_async_all();
      }
      for (int kb = 0; kb < num_kb1; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_arrive_expect_tx(fb, STAGE_BYTES);
        tma_load_2d(ring + stage * STAGE_BYTES, &q.tma_h1, fb, kb * BLOCK_K, m_blk * BLOCK_M);
        tma_load_2d(ring + stage * STAGE_BYTES + A_BYTES, &q.tma_w2, fb, kb * BLOCK_K, 0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      // W3 goes into the operand ring once every layer-2 MMA has retired (the ring is idle from then on)
      mbar_wait(smem_u32(acc1_full), 0);
      mbar_arrive_expect_tx(smem_u32(w3_full), num_kb2 * KB_BYTES);
      for (int kb = 0; kb < num_kb2; ++kb) tma_load_2d(ring + kb * KB_BYTES, &q.tma_w3, smem_u32(w3_full), kb * BLOCK_K, 0);
      pdl_launch_dependents();     // all loads issued, every CTA of the one-wave grid is resident: see gemm_kernel
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t desc_hi = (uint32_t)(desc_base() >> 32);
      const uint32_t step = (UMMA_K * 2) >> 4;      // K-major operands: 32 bytes per 16-wide K slice
      const int n2_mma = (q.N2 + 31) & ~31, n3_mma = (q.N3 + 31) & ~31;
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n2_mma >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
      const uint32_t idesc3 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n3_mma >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, accumulate = 0;
      for (int kb = 0; kb < num_kb1; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (kb == 0) trace(0x302);
        const uint32_t sa = ring + stage * STAGE_BYTES;
        const uint32_t a_lo = (1u << 16) | ((sa & 0x3FFFFu) >> 4), b_lo = (1u << 16) | (((sa + A_BYTES) & 0x3FFFFu) >> 4);
        const int k_left = q.K1 - kb * BLOCK_K;
        const int nk = k_left >= BLOCK_K ? BLOCK_K / UMMA_K : (k_left + UMMA_K - 1) / UMMA_K;
        for (int k = 0; k < nk; ++k) {
          umma_bf16(tmem_base, (uint64_t(desc_hi) << 32) | (a_lo + k * step), (uint64_t(desc_hi) << 32) | (b_lo + k * step),
                    idesc2, accumulate);
          accumulate = 1;
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(smem_u32(acc1_full));
      trace(0x303);
      // layer 3: A = the h2 tile the epilogue warps wrote to shared memory, B = W3 in the ring
      mbar_wait(smem_u32(w3_full), 0);
      mbar_wait(smem_u32(h2_ready), 0);
      tc_fence_after();
      trace(0x306);
      accumulate = 0;
      for (int kb = 0; kb < num_kb2; ++kb) {
        const uint32_t a_lo = (1u << 16) | (((h2s + kb * KB_BYTES) & 0x3FFFFu) >> 4);
        const uint32_t b_lo = (1u << 16) | (((ring + kb * KB_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          umma_bf16(tmem_base + 256, (uint64_t(desc_hi) << 32) | (a_lo + k * step), (uint64_t(desc_hi) << 32) | (b_lo + k * step),
                    idesc3, accumulate);
          accumulate = 1;
        }
      }
      umma_commit(smem_u32(acc2_full));
      // output layer: A = the h3 tile (bf16, shared memory), B = W_out as hi + mid + lo; 128 x 16 accumulator.  The three
      // terms' products are exact (bf16 x bf16), so the fp32 accumulator holds W h to fp32 round-off.
      mbar_wait(smem_u32(h3_ready), 0);
      tc_fence_after();
      const uint32_t idesc4 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(16 >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);
      const uint32_t wout = smem_u32(smem + WOUT_OFF);
      accumulate = 0;
      for (int term = 0; term < 3; ++term) {
        for (int kb = 0; kb < 2; ++kb) {
          const uint32_t a_lo = (1u << 16) | (((h2s + kb * KB_BYTES) & 0x3FFFFu) >> 4);
          const uint32_t b_lo = (1u << 16) | (((wout + term * WOUT_TERM_BYTES + kb * WOUT_KB_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_bf16(tmem_base + TMEM_OUT_COL, (uint64_t(desc_hi) << 32) | (a_lo + k * step), (uint64_t(desc_hi) << 32) | (b_lo + k * step),
                      idesc4, accumulate);
            accumulate = 1;
          }
        }
      }
      umma_commit(smem_u32(acc3_full));
    }
  } else if (warp >= 4) {
    // ================================================================ epilogues: warps 4..19
    const int quad = warp & 3, slice = (warp - 4) >> 2;
    const int et = (warp - 4) * 32 + lane;         // 0..511
    const int rt = quad * 32 + lane;               // row inside the tile
    const HgFusedHead& hd = q.head;
    const int n_out = hd.n_out;
    float* sB3 = sm;                   // [128] layer-3 bias
    float* sB = sm + 128;              // [16] output bias, [16] std, [16] log std, [16] 2 std^2
    float* sB2 = sm + 192;             // [256] layer-2 bias (a global load per use sat on the critical path of epilogue 1)
    float* sLp = reinterpret_cast<float*>(smem + LP_OFF);    // [4][128] log-prob partial sums (h2 region, dead by then)
    float* sOut = sLp + 512;                                 // [3][128 x 16] actions / mu / sigma of the tile
    // output-layer weights -> bf16 hi / mid / lo in the canonical K-major SWIZZLE_128B layout (rows >= n_out, k >= N3: zero),
    // biases and per-action constants -> shared memory, all while the main loops run
    {
      uint8_t* wb = smem + WOUT_OFF;
      for (int i = et; i < 16 * 128; i += 512) {
        const int a = i >> 7, k = i & 127;
        const float w = (a < n_out && k < q.N3) ? __ldg(hd.W + a * q.N3 + k) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        const int kb = k >> 6, kk = k & 63;
        const int off = kb * WOUT_KB_BYTES + (a >> 3) * 1024 + (a & 7) * 128 + (((kk >> 3) ^ (a & 7)) << 4) + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(wb + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(wb + WOUT_TERM_BYTES + off) = mid;
        *reinterpret_cast<__nv_bfloat16*>(wb + 2 * WOUT_TERM_BYTES + off) = lo;
      }
      if (et < 128) sB3[et] = (q.b3 && et < q.N3) ? __ldg(q.b3 + et) : 0.f;
      if (et >= 256) sB2[et - 256] = (et - 256 < q.N2) ? __ldg(q.b2 + et - 256) : 0.f;
      if (et < 16) {
        const float sd = (hd.kind == HG_HEAD_POLICY && et < n_out) ? __ldg(hd.std + et) : 1.f;
        sB[et] = et < n_out ? __ldg(hd.b + et) : 0.f;
        sB[16 + et] = sd;
        sB[32 + et] = logf(sd);
        sB[48 + et] = 2.f * sd * sd;
      }
    }
    // The four normals of this thread's outputs (row, 4 * slice ..): they depend on (seed, counter, row, slice) only, so
    // they are drawn NOW, behind the layer-2 main loop, instead of on the critical path after the output layer (where the
    // Philox rounds + Box-Muller took ~2 us of the actor tile's head)
    const int row = m_blk * BLOCK_M + rt;
    const int a_lo = slice * 4;
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    if (hd.kind == HG_HEAD_POLICY && P.eps == nullptr && a_lo < n_out) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)slice, (uint32_t)P.counter, (uint32_t)(P.counter >> 32)),
                                    make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
      const float2 n01 = box_muller(r.x, r.y), n23 = box_muller(r.z, r.w);
      e[0] = n01.x; e[1] = n01.y; e[2] = n23.x; e[3] = n23.y;
    }
    epi_bar_sync();                                                 // sB2 / sB3 / sB staged by every epilogue thread
    // ---- epilogue 1: h2 = bf16(elu(acc + b2)) -> shared memory, K-major SWIZZLE_128B (16-byte chunk index XOR row % 8)
    mbar_wait(smem_u32(acc1_full), 0);
    tc_fence_after();
    if (warp == 4 && lane == 0) trace(0x304);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int n0 = slice * 64 + c * 32;
      if (n0 < q.N2) {                                     // N2 % 64 == 0: a live chunk is a full chunk
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + n0, v);
        tmem_ld_wait();
        const uint32_t row_addr = h2s + (n0 >> 6) * KB_BYTES + (rt >> 3) * 1024 + (rt & 7) * 128;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = __uint_as_float(v[p * 8 + j]) + sB2[n0 + p * 8 + j];
            z[j] = t > 0.f ? t : exp2f_fast(t * 1.4426950408889634f) - 1.f;
          }
          const uint4 w = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
          const int chunk = ((n0 & 63) >> 3) + p;          // 16-byte chunk of the row's 128-byte k-block line
          sts128(row_addr + ((chunk ^ (rt & 7)) << 4), w);
        }
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(h2_ready));
    if (warp == 4 && lane == 0) trace(0x305);

    // ---- epilogue 2: h3 = bf16(elu(acc + b3)) -> shared memory (over the h2 tile: layer 3 has read it), the A operand of the
    // output layer; dead 32-column chunks (N3 < 128) are zero-filled, their weights are zero too
    mbar_wait(smem_u32(acc2_full), 0);
    tc_fence_after();
    if (warp == 4 && lane == 0) trace(0x307, g);
    {
      const int n0 = slice * 32;
      uint32_t v[32];
      const bool live = n0 < q.N3;                  // N3 % 32 == 0
      if (live) {
        tmem_ld_32x32(tmem_base + (uint32_t(quad * 32) << 16) + 256 + n0, v);
        tmem_ld_wait();
      }
      const uint32_t row_addr = h2s + (n0 >> 6) * KB_BYTES + (rt >> 3) * 1024 + (rt & 7) * 128;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        uint4 w = make_uint4(0, 0, 0, 0);
        if (live) {
          float z[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = __uint_as_float(v[p * 8 + j]) + sB3[n0 + p * 8 + j];
            z[j] = x > 0.f ? x : exp2f_fast(x * 1.4426950408889634f) - 1.f;
          }
          w = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]), pack_bf16x2(z[6], z[7]));
        }
        const int chunk = ((n0 & 63) >> 3) + p;
        sts128(row_addr + ((chunk ^ (rt & 7)) << 4), w);
      }
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(h3_ready));
    if (warp == 4 && lane == 0) trace(0x309, g);

    // ---- epilogue 3: the head.  Thread (row, slice) finishes outputs 4 * slice .. 4 * slice + 3 of its row: it reads just
    // those four accumulator columns, and one Philox call (keyed by (row, slice)) gives their four normals.
    mbar_wait(smem_u32(acc3_full), 0);
    tc_fence_after();
    float lp = 0.f;
    if (a_lo < n_out) {                                     // (warp-uniform)
      uint32_t o4[4];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(o4[0]), "=r"(o4[1]), "=r"(o4[2]), "=r"(o4[3])
                   : "r"(tmem_base + (uint32_t(quad * 32) << 16) + TMEM_OUT_COL + a_lo) : "memory");
      tmem_ld_wait();
      if (row < q.M) {
        if (hd.kind == HG_HEAD_POLICY) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int a = a_lo + k;
            if (a < n_out) {
              const float m = __uint_as_float(o4[k]) + sB[a], sd = sB[16 + a];
              const float ee = P.eps ? __ldg(P.eps + (long long)row * n_out + a) : e[k];
              const float act = m + sd * ee;                 // Normal(mu, std).sample()  (actor_critic.py:129,133)
              const float d = act - m;
              lp += -(d * d) / sB[48 + a] - sB[32 + a] - LOG_SQRT_2PI;   // Normal.log_prob
              sOut[rt * n_out + a] = act;                    // staged: the tile's [128][n_out] block is contiguous in global
              sOut[2048 + rt * n_out + a] = m;               // memory, written below with coalesced 16-byte stores (12 scattered
              sOut[4096 + rt * n_out + a] = sd;              // 4-byte stores per thread cost ~2 us of LSU sector traffic)
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (a_lo + k < n_out) hd.out[(long long)row * n_out + a_lo + k] = __uint_as_float(o4[k]) + sB[a_lo + k];
        }
      }
    }
    tc_fence_before();
    if (warp == 4 && lane == 0) trace(0x30a, g);
    if (hd.kind == HG_HEAD_POLICY) {       // log-prob = fixed-order sum of the four slices' terms
      sLp[slice * 128 + rt] = lp;
      epi_bar_sync();
      if (slice == 0 && row < q.M) hd.log_prob[row] = ((sLp[rt] + sLp[128 + rt]) + sLp[256 + rt]) + sLp[384 + rt];
      const int count = min(BLOCK_M, q.M - m_blk * BLOCK_M) * n_out;
      const long long g0 = (long long)m_blk * BLOCK_M * n_out;
#pragma unroll 1
      for (int arr = 0; arr < 3; ++arr) {
        float* dst = (arr == 0 ? hd.actions : arr == 1 ? hd.out : hd.sigma) + g0;
        const float* src = sOut + arr * 2048;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (count & 3) == 0) {
          for (int i = et; i < (count >> 2); i += 512) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
        } else {
          for (int i = et; i < count; i += 512) dst[i] = src[i];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace(0x308, g);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<tail::TMEM_COLS>(tmem_base);
  }
}

__global__ void __launch_bounds__(tail::THREADS, 1) rollout_tail_kernel(const __grid_constant__ TailParams P) {
  const int g = blockIdx.x / P.tiles_m;
  tail_body(P, g, blockIdx.x - g * P.tiles_m, nullptr, 0u);
}

// ------------------------------------------------------------------------------------------ fused rollout step
// Layer 1 AND the tail (layer 2 + layer 3 + output layers + sampling) of one env step in ONE launch: the first `num_l1`
// CTAs run the grouped layer-1 GEMM (m-major tile order: the row blocks complete front to back), the remaining CTAs each
// own one (128-row block, net) of the tail and start as soon as the layer-1 tiles of THEIR row block are published.
// Removes one kernel boundary per step (~2 us: completion + flush + dependent launch) and overlaps the tail's main loops
// and epilogues with the rest of layer 1; both parts pull their operands through the same chip-wide L2 -> SM cap, which the
// layer-1 CTAs alone (fill-bound at ~42 B/clk/SM on 148 SMs) do not need all SMs to saturate.
struct alignas(64) StepParams {
  GroupParams gp;
  TailParams tp;
  unsigned* flags;       // [nets * tiles_m] tile counters, monotonically increasing over the launches of the plan
  unsigned target[2];    // per net: counter value at which all layer-1 tiles of a row block are in
  int num_l1;            // CTAs of the layer-1 part
};
static_assert(Cfg<128, HG_EPI_BIAS_ELU_BF16, 1>::THREADS == tail::THREADS, "one block shape for both parts");

__global__ void __launch_bounds__(tail::THREADS, 1) rollout_step_kernel(const __grid_constant__ StepParams S) {
  if ((int)blockIdx.x < S.num_l1) {
    gemm_body<128, HG_EPI_BIAS_ELU_BF16, 1, true>(S.gp, (int)blockIdx.x, S.num_l1, S.flags);
  } else {
    const int j = (int)blockIdx.x - S.num_l1;
    const int g = j % S.tp.num_nets;
    tail_body(S.tp, g, j / S.tp.num_nets, S.flags, S.target[g]);
  }
}

HG_TRACE_SETTER(trace_set_gemm)

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over row-major [outer, inner] with row stride ld elements, box {64, box_outer}, 128 B swizzle.
// 2-D bf16 tensor map over the row-major output [outer, inner]: box {32 columns, 32 rows} = one warp's chunk, 64 B swizzle.
static int make_tmap_out(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld) {
  EncodeTiledFn enc = get_encode_fn();
  HG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (output) failed with %d (inner %lld outer %lld ld %lld)", (int)r, inner, outer, ld);
  return HG_OK;
}

static int make_tmap(CUtensorMap* m, const void* base, long long inner, long long outer, long long ld, int box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  HG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  HG_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base address must be 16-byte aligned");
  HG_CHECK(ld % 8 == 0, "leading dimension %lld must be a multiple of 8 bf16 elements", ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (inner %lld outer %lld ld %lld)", (int)r, inner,
           outer, ld);
  return HG_OK;
}

template <int BLOCK_N, int EPI, int CG>
static int launch(const GroupParams& gp, cudaStream_t stream) {
  const int smem = Cfg<BLOCK_N, EPI, CG>::SMEM_TOTAL;
  static bool configured = false;
  if (!configured) {
    HG_CUDA(cudaFuncSetAttribute(gemm_kernel<BLOCK_N, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int max_groups = num_sms() / CG;
  const int groups = gp.total_tiles < max_groups ? gp.total_tiles : max_groups;
  HG_CUDA(launch_kernel(gemm_kernel<BLOCK_N, EPI, CG>, dim3(groups * CG), dim3(Cfg<BLOCK_N, EPI, CG>::THREADS), smem, stream, CG, gp));
  HG_LAUNCH_CHECK();
  count_launch();
  return HG_OK;
}

template <int BLOCK_N, int CG>
static int dispatch_epi(const GroupParams& gp, int epi, cudaStream_t s) {
  switch (epi) {
    case HG_EPI_STORE_F32: return launch<BLOCK_N, HG_EPI_STORE_F32, CG>(gp, s);
    case HG_EPI_BIAS_ELU_BF16: return launch<BLOCK_N, HG_EPI_BIAS_ELU_BF16, CG>(gp, s);
    case HG_EPI_DELU_BF16: return launch<BLOCK_N, HG_EPI_DELU_BF16, CG>(gp, s);
    case HG_EPI_ATOMIC_F32: return launch<BLOCK_N, HG_EPI_ATOMIC_F32, CG>(gp, s);
    case HG_EPI_BIAS_BF16: return launch<BLOCK_N, HG_EPI_BIAS_BF16, CG>(gp, s);
  }
  set_error("unknown epilogue %d", epi);
  return HG_ERR_INVALID;
}

static int dispatch(const GroupParams& gp, int epi, int block_n, int cg, cudaStream_t s) {
  if (epi == HG_EPI_BIAS_ELU_HEAD) return launch<128, HG_EPI_BIAS_ELU_HEAD, 1>(gp, s);   // validated: block_n 128, cg 1
  if (cg == 2) return block_n == 128 ? dispatch_epi<128, 2>(gp, epi, s) : dispatch_epi<256, 2>(gp, epi, s);
  return block_n == 128 ? dispatch_epi<128, 1>(gp, epi, s) : dispatch_epi<256, 1>(gp, epi, s);
}

}  // namespace hg

namespace hg {
// Validate the problems, encode the TMA tensor maps and lay out the tile schedule (host side, ~1 us per tensor map).
static int build_params(const HgGemmProblem* pr, int n, int epilogue, int block_n, int cg, GroupParams& gp) {
  HG_CHECK(cg == 1 || cg == 2, "cta_group must be 1 or 2");
  HG_CHECK(pr != nullptr && n >= 1 && n <= HG_MAX_GROUP, "num_problems must be in [1,%d]", HG_MAX_GROUP);
  HG_CHECK(block_n == 128 || block_n == 256, "block_n must be 128 or 256");
  HG_CHECK(epilogue >= HG_EPI_STORE_F32 && epilogue <= HG_EPI_BIAS_ELU_HEAD, "unknown epilogue %d", epilogue);
  const bool fused_head = epilogue == HG_EPI_BIAS_ELU_HEAD;
  HG_CHECK(!fused_head || (block_n == 128 && cg == 1), "the fused-head epilogue needs block_n 128 and cta_group 1");
  memset(&gp, 0, sizeof(gp));
  gp.num_problems = n;
  int tiles = 0, colsum_cols = 0;
  for (int g = 0; g < n; ++g) {
    const HgGemmProblem& h = pr[g];
    DevProblem& q = gp.p[g];
    HG_CHECK(h.M > 0 && h.N > 0 && h.K > 0, "problem %d: empty dims", g);
    HG_CHECK(h.A && h.B && (h.out || fused_head), "problem %d: null pointer", g);
    HG_CHECK(!fused_head || (h.N <= 128 && h.N % 32 == 0), "problem %d: the fused head needs N <= 128, N %% 32 == 0", g);
    HG_CHECK(h.k_splits >= 1, "problem %d: k_splits", g);
    HG_CHECK(h.k_splits == 1 || epilogue == HG_EPI_ATOMIC_F32, "split-K needs the atomic epilogue");
    HG_CHECK(epilogue != HG_EPI_DELU_BF16 || h.act, "problem %d: DELU epilogue needs act", g);
    q.M = h.M; q.N = h.N; q.K = h.K;
    q.a_mn = h.a_mn_major ? 1 : 0;
    q.b_mn = h.b_mn_major ? 1 : 0;
    q.num_kb = (h.K + BLOCK_K - 1) / BLOCK_K;
    int ks = h.k_splits < q.num_kb ? h.k_splits : q.num_kb;
    q.kb_per_split = (q.num_kb + ks - 1) / ks;
    q.k_splits = (q.num_kb + q.kb_per_split - 1) / q.kb_per_split;   // no empty split
    q.tiles_m = (h.M + BLOCK_M * cg - 1) / (BLOCK_M * cg);
    q.tiles_n = (h.N + block_n - 1) / block_n;
    q.inv_tiles_m = (unsigned)((0x100000000ull + (unsigned)q.tiles_m - 1) / (unsigned)q.tiles_m);
    q.inv_tiles_n = (unsigned)((0x100000000ull + (unsigned)q.tiles_n - 1) / (unsigned)q.tiles_n);
    HG_CHECK((unsigned long long)q.tiles_m * q.tiles_n * q.k_splits * (unsigned long long)(q.tiles_m > q.tiles_n ? q.tiles_m : q.tiles_n) < 0x100000000ull,
             "problem %d: too many tiles for the reciprocal tile decode", g);
    q.tile_begin = tiles;
    tiles += q.tiles_m * q.tiles_n * q.k_splits;
    q.tile_end = tiles;
    q.out = h.out; q.ldo = h.ldo; q.bias = h.bias;
    q.act = reinterpret_cast<const __nv_bfloat16*>(h.act); q.ldact = h.ldact; q.colsum = h.colsum;
    q.colsum_off = colsum_cols;
    HG_CHECK((h.out_frame_src == 0 && h.out_frame_dst == 0) ||
                 (epilogue == HG_EPI_ATOMIC_F32 && h.out_frame_dst >= 1 && h.out_frame_src >= h.out_frame_dst),
             "problem %d: out_frame_src / out_frame_dst need the atomic epilogue and src >= dst >= 1", g);
    q.out_fs = h.out_frame_src; q.out_fd = h.out_frame_dst;
    if (epilogue == HG_EPI_DELU_BF16 && h.colsum) colsum_cols += (h.N + 31) & ~31;
    HG_CHECK(colsum_cols <= HG_MAX_COLSUM_COLS, "problem %d: a DELU launch can column-sum at most %d columns in total", g, HG_MAX_COLSUM_COLS);
    if (epilogue == HG_EPI_BIAS_ELU_BF16 || epilogue == HG_EPI_DELU_BF16 || epilogue == HG_EPI_BIAS_BF16)
      HG_CHECK(h.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(h.out) & 15) == 0, "problem %d: bf16 output alignment", g);
    if (epilogue == HG_EPI_DELU_BF16)
      HG_CHECK(h.ldact % 8 == 0 && (reinterpret_cast<uintptr_t>(h.act) & 15) == 0, "problem %d: act alignment", g);
    int rc;
    if (!q.a_mn) rc = make_tmap(&gp.tma_a[g], h.A, h.K, h.M, h.lda, BLOCK_M);
    else rc = make_tmap(&gp.tma_a[g], h.A, h.M, h.K, h.lda, BLOCK_K);
    if (rc != HG_OK) return rc;
    if (!q.b_mn) rc = make_tmap(&gp.tma_b[g], h.B, h.K, h.N, h.ldb, block_n / cg);
    else rc = make_tmap(&gp.tma_b[g], h.B, h.N, h.K, h.ldb, BLOCK_K);
    if (rc != HG_OK) return rc;
    if (epilogue == HG_EPI_BIAS_ELU_BF16 || epilogue == HG_EPI_DELU_BF16 || epilogue == HG_EPI_BIAS_BF16) {
      rc = make_tmap_out(&gp.tma_out[g], h.out, h.N, h.M, h.ldo);
      if (rc != HG_OK) return rc;
    }
  }
  gp.total_tiles = tiles;
  gp.m_major = n > 1 ? 1 : 0;
  gp.tiles_per_m = 0;
  for (int g = 0; g < n; ++g) {
    if (gp.p[g].tiles_m != gp.p[0].tiles_m || gp.p[g].k_splits != 1) gp.m_major = 0;
    gp.tiles_per_m += gp.p[g].tiles_n;
  }
  // Measured on B200 (tools/phase_times.py, 4096 envs x 60): m-major leaves the update unchanged (5.29 vs 5.29 ms: the
  // second read of obs was not the limiter) and slows the rollout launches (3.52 vs 3.38 ms), so problem-major stays the
  // default; HG_TILE_ORDER=m selects m-major for experiments.
  const char* order = getenv("HG_TILE_ORDER");
  if (!(order && order[0] == 'm')) gp.m_major = 0;
  return HG_OK;
}
}  // namespace hg

struct HgGemmPlan {
  hg::GroupParams gp;
  int epilogue, block_n, cta_group;
  bool tail = false;          // hg_gemm_plan_create_tail: layer 2 + layer 3 + head in one launch
  hg::TailParams tp;
  bool step = false;          // hg_gemm_plan_create_step: layer 1 + tail in one launch (gp = layer 1, tp = tail)
  unsigned* step_flags = nullptr;
  mutable unsigned step_epoch = 0;
  int step_num_l1 = 0;
  ~HgGemmPlan() {
    if (step_flags) cudaFree(step_flags);
  }
};

extern "C" int hg_gemm_grouped(const HgGemmProblem* pr, int n, int epilogue, int block_n, int cta_group, void* stream) {
  using namespace hg;
  HG_CHECK(epilogue != HG_EPI_BIAS_ELU_HEAD, "HG_EPI_BIAS_ELU_HEAD launches go through hg_gemm_plan_create_heads");
  GroupParams gp;
  int rc = build_params(pr, n, epilogue, block_n, cta_group, gp);
  if (rc != HG_OK) return rc;
  return dispatch(gp, epilogue, block_n, cta_group, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int hg_gemm_plan_create(const HgGemmProblem* pr, int n, int epilogue, int block_n, int cta_group, HgGemmPlan** out) {
  using namespace hg;
  HG_CHECK(out != nullptr, "hg_gemm_plan_create: null output");
  HG_CHECK(epilogue != HG_EPI_BIAS_ELU_HEAD, "HG_EPI_BIAS_ELU_HEAD plans are made by hg_gemm_plan_create_heads");
  HgGemmPlan* plan = new HgGemmPlan();
  int rc = build_params(pr, n, epilogue, block_n, cta_group, plan->gp);
  if (rc != HG_OK) {
    delete plan;
    return rc;
  }
  plan->epilogue = epilogue;
  plan->block_n = block_n;
  plan->cta_group = cta_group;
  *out = plan;
  return HG_OK;
}

extern "C" int hg_gemm_plan_run(const HgGemmPlan* plan, void* stream) {
  using namespace hg;
  HG_CHECK(plan != nullptr, "hg_gemm_plan_run: null plan");
  HG_CHECK(plan->epilogue != HG_EPI_BIAS_ELU_HEAD, "hg_gemm_plan_run: fused-head plans run through hg_gemm_plan_run_heads");
  return dispatch(plan->gp, plan->epilogue, plan->block_n, plan->cta_group, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" void hg_gemm_plan_destroy(HgGemmPlan* plan) { delete plan; }

extern "C" int hg_gemm_plan_create_heads(const HgGemmProblem* pr, const HgFusedHead* heads, int n, HgGemmPlan** out) {
  using namespace hg;
  HG_CHECK(out != nullptr && heads != nullptr, "hg_gemm_plan_create_heads: null argument");
  static_assert(HG_EPI_WARPS == 16, "the fused-head epilogue synchronises 16 epilogue warps");
  HgGemmPlan* plan = new HgGemmPlan();
  int rc = build_params(pr, n, HG_EPI_BIAS_ELU_HEAD, 128, 1, plan->gp);
  for (int g = 0; rc == HG_OK && g < n; ++g) {
    const HgFusedHead& h = heads[g];
    const bool ok = (h.kind == HG_HEAD_POLICY || h.kind == HG_HEAD_LINEAR) && h.n_out >= 1 && h.n_out <= 16 && h.W && h.b && h.out &&
                    (h.kind != HG_HEAD_POLICY || (h.std && h.actions && h.sigma && h.log_prob));
    if (!ok) {
      set_error("hg_gemm_plan_create_heads: bad head %d", g);
      rc = HG_ERR_INVALID;
    }
    plan->gp.heads[g] = h;
  }
  if (rc != HG_OK) {
    delete plan;
    return rc;
  }
  plan->epilogue = HG_EPI_BIAS_ELU_HEAD;
  plan->block_n = 128;
  plan->cta_group = 1;
  *out = plan;
  return HG_OK;
}

namespace hg {
static int fill_tail_params(const HgGemmProblem* l2, const HgGemmProblem* l3, const HgFusedHead* heads, int n, TailParams& tp) {
  memset(&tp, 0, sizeof(tp));
  tp.num_nets = n;
  for (int g = 0; g < n; ++g) {
    const HgGemmProblem& a = l2[g];
    const HgGemmProblem& b = l3[g];
    const HgFusedHead& h = heads[g];
    const bool ok = a.A && a.B && b.B && a.bias && a.M > 0 && a.K > 0 && !a.a_mn_major && !a.b_mn_major && !b.b_mn_major &&
                    a.N >= 64 && a.N <= 256 && a.N % 64 == 0 && b.K == a.N && b.N >= 32 && b.N <= 128 && b.N % 32 == 0 &&
                    a.M == l2[0].M && (h.kind == HG_HEAD_POLICY || h.kind == HG_HEAD_LINEAR) && h.n_out >= 1 && h.n_out <= 16 &&
                    h.W && h.b && h.out && (h.kind != HG_HEAD_POLICY || (h.std && h.actions && h.sigma && h.log_prob));
    if (!ok) {
      set_error("rollout tail: net %d: needs K-major operands, 64 <= N2 <= 256 (multiple of 64), K3 == N2, "
                "32 <= N3 <= 128 (multiple of 32), equal M, a valid head", g);
      return HG_ERR_INVALID;
    }
    TailNet& q = tp.net[g];
    q.M = a.M; q.K1 = a.K; q.N2 = a.N; q.N3 = b.N;
    q.b2 = a.bias; q.b3 = b.bias; q.head = h;
    int rc = make_tmap(&q.tma_h1, a.A, a.K, a.M, a.lda, BLOCK_M);
    if (rc == HG_OK) rc = make_tmap(&q.tma_w2, a.B, a.K, a.N, a.ldb, 256);
    if (rc == HG_OK) rc = make_tmap(&q.tma_w3, b.B, b.K, b.N, b.ldb, 128);
    if (rc != HG_OK) return rc;
  }
  tp.tiles_m = (l2[0].M + BLOCK_M - 1) / BLOCK_M;
  return HG_OK;
}
}  // namespace hg

extern "C" int hg_gemm_plan_create_tail(const HgGemmProblem* l2, const HgGemmProblem* l3, const HgFusedHead* heads, int n,
                                        HgGemmPlan** out) {
  using namespace hg;
  HG_CHECK(out && l2 && l3 && heads && n >= 1 && n <= 2, "hg_gemm_plan_create_tail: 1 or 2 nets, non-null arguments");
  HgGemmPlan* plan = new HgGemmPlan();
  plan->tail = true;
  plan->epilogue = HG_EPI_BIAS_ELU_HEAD;
  plan->block_n = 128;
  plan->cta_group = 1;
  const int rc = fill_tail_params(l2, l3, heads, n, plan->tp);
  if (rc != HG_OK) {
    delete plan;
    return rc;
  }
  *out = plan;
  return HG_OK;
}

extern "C" int hg_gemm_plan_create_step(const HgGemmProblem* l1, const HgGemmProblem* l2, const HgGemmProblem* l3,
                                        const HgFusedHead* heads, int n, HgGemmPlan** out) {
  using namespace hg;
  HG_CHECK(out && l1 && l2 && l3 && heads && n >= 1 && n <= 2, "hg_gemm_plan_create_step: 1 or 2 nets, non-null arguments");
  HgGemmPlan* plan = new HgGemmPlan();
  plan->tail = true;
  plan->step = true;
  plan->epilogue = HG_EPI_BIAS_ELU_HEAD;
  plan->block_n = 128;
  plan->cta_group = 1;
  int rc = build_params(l1, n, HG_EPI_BIAS_ELU_BF16, 128, 1, plan->gp);
  if (rc == HG_OK) rc = fill_tail_params(l2, l3, heads, n, plan->tp);
  for (int g = 0; rc == HG_OK && g < n; ++g) {
    // layer 1 of net g writes exactly what the tail of net g reads, whole 32-column chunks (every store a bulk store)
    const bool ok = l1[g].M == l2[g].M && l1[g].out == l2[g].A && l1[g].ldo == l2[g].lda && l1[g].N == l2[g].K && l1[g].N % 32 == 0 &&
                    l1[g].k_splits == 1 && !l1[g].a_mn_major && !l1[g].b_mn_major;
    if (!ok) {
      set_error("hg_gemm_plan_create_step: net %d: layer 1 must produce the tail's h1 (same M, buffer, pitch; N %% 32 == 0)", g);
      rc = HG_ERR_INVALID;
    }
  }
  if (rc == HG_OK) {
    const int tail_ctas = n * plan->tp.tiles_m;
    plan->step_num_l1 = num_sms() - tail_ctas;
    if (plan->step_num_l1 < 32) {
      set_error("hg_gemm_plan_create_step: %d tail CTAs leave fewer than 32 of %d SMs for layer 1", tail_ctas, num_sms());
      rc = HG_ERR_INVALID;
    }
  }
  if (rc == HG_OK) {
    plan->gp.m_major = 1;        // row blocks complete front to back (all problems have tiles_m row blocks, no split-K)
    const size_t bytes = sizeof(unsigned) * n * plan->tp.tiles_m;
    if (cudaMalloc(&plan->step_flags, bytes) != cudaSuccess || cudaMemset(plan->step_flags, 0, bytes) != cudaSuccess) {
      set_error("hg_gemm_plan_create_step: cannot allocate the tile counters");
      rc = HG_ERR_CUDA;
    }
  }
  if (rc != HG_OK) {
    delete plan;
    return rc;
  }
  *out = plan;
  return HG_OK;
}

extern "C" int hg_gemm_plan_run_heads(const HgGemmPlan* plan, const float* eps, uint64_t seed, uint64_t counter, void* stream) {
  using namespace hg;
  HG_CHECK(plan != nullptr && plan->epilogue == HG_EPI_BIAS_ELU_HEAD, "hg_gemm_plan_run_heads: not a fused-head plan");
  if (plan->step) {
    constexpr int SMEM = Cfg<128, HG_EPI_BIAS_ELU_BF16, 1>::SMEM_TOTAL > tail::SMEM_TOTAL ? Cfg<128, HG_EPI_BIAS_ELU_BF16, 1>::SMEM_TOTAL
                                                                                          : tail::SMEM_TOTAL;
    static bool configured = false;
    if (!configured) {
      HG_CUDA(cudaFuncSetAttribute(rollout_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      configured = true;
    }
    StepParams sp;
    sp.gp = plan->gp;
    sp.tp = plan->tp;
    sp.tp.eps = eps;
    sp.tp.seed = seed;
    sp.tp.counter = counter;
    sp.flags = plan->step_flags;
    sp.num_l1 = plan->step_num_l1;
    const unsigned epoch = ++plan->step_epoch;
    for (int g = 0; g < sp.tp.num_nets; ++g) sp.target[g] = epoch * (unsigned)(plan->gp.p[g].tiles_n * Cfg<128, HG_EPI_BIAS_ELU_BF16, 1>::EPI_WARPS);
    HG_CUDA(launch_kernel(rollout_step_kernel, dim3(sp.num_l1 + sp.tp.num_nets * sp.tp.tiles_m), dim3(tail::THREADS), SMEM,
                          reinterpret_cast<cudaStream_t>(stream), 1, sp));
    HG_LAUNCH_CHECK();
    count_launch();
    return HG_OK;
  }
  if (plan->tail) {
    static bool configured = false;
    if (!configured) {
      HG_CUDA(cudaFuncSetAttribute(rollout_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tail::SMEM_TOTAL));
      configured = true;
    }
    TailParams tp = plan->tp;
    tp.eps = eps;
    tp.seed = seed;
    tp.counter = counter;
    HG_CUDA(launch_kernel(rollout_tail_kernel, dim3(tp.num_nets * tp.tiles_m), dim3(tail::THREADS), tail::SMEM_TOTAL,
                          reinterpret_cast<cudaStream_t>(stream), 1, tp));
    HG_LAUNCH_CHECK();
    count_launch();
    return HG_OK;
  }
  GroupParams gp = plan->gp;
  gp.eps = eps;
  gp.seed = seed;
  gp.counter = counter;
  return dispatch(gp, HG_EPI_BIAS_ELU_HEAD, 128, 1, reinterpret_cast<cudaStream_t>(stream));
}

// The rollout forward of one env step in ONE call: `n` plain plans (layer 1, ...) followed by the fused-head plan.
// Same launches as hg_gemm_plan_run x n + hg_gemm_plan_run_heads; saves the per-call host overhead of the binding.
extern "C" int hg_rollout_forward(const HgGemmPlan* const* plans, int n, const HgGemmPlan* head_plan, const float* eps,
                                  uint64_t seed, uint64_t counter, void* stream) {
  using namespace hg;
  HG_CHECK((plans != nullptr || n == 0) && n >= 0 && head_plan != nullptr, "hg_rollout_forward: bad arguments");
  for (int i = 0; i < n; ++i) {
    int rc = hg_gemm_plan_run(plans[i], stream);
    if (rc != HG_OK) return rc;
  }
  return hg_gemm_plan_run_heads(head_plan, eps, seed, counter, stream);
}
