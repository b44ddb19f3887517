// lce_b200_pw.cuh -- the fp32 pointwise (1x1, stride 1) CONV_2D either side of the binary path,
// on the tensor cores: out[M][N] = act(A[M][K] * W[N][K]^T + bias[N])  (TFLite reference
// semantics: tensorflow/lite/kernels/internal/reference/conv.h:27; the QuickNet transition and
// stem pointwise layers and Bi-RealNet's shortcut convolutions).
//
// An fp32 FMA GEMM of these shapes is issue-bound at 4-5x its HBM time. tcgen05.mma kind::tf32
// reads fp32 operands from shared memory and TRUNCATES them to tf32 (measured: tools/tc_probe.cu
// T7), so the kernel splits every operand once in shared memory,
//     hi = rna_tf32(x),  lo = rna_tf32(x - hi)                (x - hi is exact in fp32)
// and accumulates hi*hi + hi*lo + lo*hi in fp32 TMEM: the dropped terms are ~2^-22 |a||w| with
// random sign, the same order as the rounding of an fp32 FMA chain of this length (probe T7
// reports the measured error). Three tensor passes still leave the kernel HBM-bound.
//
// Per CTA (persistent, one per SM), items = (128-row tile, 128-column tile):
//   warp 20     TMA: A[128 x 32 floats] (+ W[128 x 32]) per K block, SWIZZLE_128B, 3 stages
//   warps 0-3   row = lane: read the A row, write hi | lo to TENSOR MEMORY (tcgen05.st, 64 columns
//               per stage) -- the A operand then costs no shared-memory bandwidth in the MMA;
//               split the W tile hi/lo in place (elementwise, so the swizzle needs no decoding)
//   warp 21     12 x tcgen05.mma (A from TMEM, W from shared memory; M 128, N 128, K 8) per K
//               block into one of 2 TMEM accumulators
//   warps 4-19  four column groups of 32: tcgen05.ld -> + bias, activation (+ LceQuantize bits)
//               -> swizzled staging -> TMA store (one 32 x 32 box per warp and item)
// (the single-thread roles sit at the highest warp ids: the warp scheduler favours them)
// K = 16 (QuickNet's 16 -> 64 pointwise) runs as pixel PAIRS: A'[M/2][32] is the same memory,
// W' = diag(W, W) is built in shared memory once, out'[M/2][128] is the same memory as out.
#ifndef LCE_B200_PW_CUH_
#define LCE_B200_PW_CUH_

#include "lce_b200_tc.cuh"

namespace lce {
namespace pw {

using namespace lce::tc;

constexpr int kPwThreads = 704;           // 4 split + 16 epilogue + producer + MMA warps
constexpr int kPwNS = 3;                  // A / W stages
constexpr int kPwTile = 16384;            // 128 rows x 128 B
constexpr int kPwWStage = 2 * kPwTile;    // W: hi | lo
constexpr int kPwNAcc = 2;                // TMEM accumulators of 128 columns (columns 0..255)
constexpr int kPwACol = 256;              // A stages in TMEM: 64 columns each (hi 32 | lo 32)
constexpr int kPwEpiBytes = 16 * 4096;    // 16 warps x (32 rows x 128 B)
constexpr int kPwBarBytes = 1024;
constexpr size_t kPwSmem = kPwNS * (kPwTile + kPwWStage) + kPwEpiBytes + kPwBarBytes + 1024;

struct PwParams {
  long long M;          // rows (pixel pairs in pairs mode)
  int N, KB;            // columns (multiple of 128), K blocks of 32 floats
  int n_tiles, m_tiles;
  int act;              // LCE_ACT_*
  int pairs;            // K = 16, N = 64 run as [M/2][32] x diag(W, W)
  int w_resident;       // n_tiles == 1 and KB <= kPwNS: the weights are staged and split once
  int bias_mask;        // column -> bias index (pairs: & 63)
  const float* filter;  // [N][K] (read directly only in pairs mode)
  const float* bias;    // may be null
  int32_t* packed;      // optional LceQuantize of the output, [M][N / 32]
  long long* prof;      // optional [22 warps][8] cycle counters of block 0 (LCE_TC_PROF build)
};

__device__ __forceinline__ void mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc),
               "r"(acc) : "memory");
}
// K-major SWIZZLE_128B operand: 8-row atoms of 1024 B (SBO), version 1, layout type 2
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kSdescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
// D = f32, A = B = tf32, K-major, N = 128, M = 128
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest (ties away from zero) to tf32, as cvt.rna.tf32.f32 does -- but on the integer
// ALU at full rate: the conversion instruction issues at a fraction of it, and 64 of them per
// thread and K block made the split warps the bottleneck (profiles/r02_pw_roles.txt).
// (No Inf / NaN handling: a value within 2^-11 of FLT_MAX would round to Inf.)
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split4(float4 x, float4* hi, float4* lo) {
  float4 h = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
  *hi = h;
  *lo = make_float4(to_tf32(x.x - h.x), to_tf32(x.y - h.y), to_tf32(x.z - h.z), to_tf32(x.w - h.w));
}
// split one 16 KB tile in place (hi) and into the tile 16 KB above it (lo); 128 threads. All
// loads first: the in-place stores alias the loads, and the compiler would otherwise serialise
// load -> convert -> store eight times over.
__device__ __forceinline__ void split_tile(unsigned char* tile, int t) {
  float4* c = reinterpret_cast<float4*>(tile) + t;
  float4 x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = c[128 * j];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 h, l;
    split4(x[j], &h, &l);
    c[128 * j] = h;
    c[128 * j + kPwTile / 16] = l;
  }
}

__global__ void __launch_bounds__(kPwThreads, 1)
pw_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_out, const PwParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* a_st = smem;                               // [kPwNS] raw A tiles
  unsigned char* w_st = a_st + kPwNS * kPwTile;             // [kPwNS][hi | lo]
  unsigned char* epi = w_st + kPwNS * kPwWStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kPwEpiBytes);
  uint64_t* full = bars;                  // TMA landed                          (1 + tx)
  uint64_t* ready = bars + kPwNS;         // A in TMEM, W hi / lo written        (128)
  uint64_t* empty = bars + 2 * kPwNS;     // MMAs of the stage retired           (commit)
  uint64_t* acc_full = bars + 3 * kPwNS;
  uint64_t* acc_empty = acc_full + kPwNAcc;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(acc_empty + kPwNAcc);

  const int tid = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kPwNS; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&ready[i], 128);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kPwNAcc; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 16);
    }
    fence_barrier_init();
  }
  if (wid == 20) tmem_alloc(tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int items = p.n_tiles * p.m_tiles;
  const bool prof = LCE_TC_PROF != 0 && p.prof != nullptr && blockIdx.x == 0 && lane == 0;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = prof ? clock64() : 0;

  if (wid == 20) {
    // ------------------------------------------------ TMA producer
    uint32_t cnt = 0;
    bool first = true;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int mt = it / p.n_tiles, nt = it - mt * p.n_tiles;
      for (int kb = 0; kb < p.KB; ++kb, ++cnt) {
        const int s = cnt % kPwNS;
        mbar_wait_prof(&empty[s], ((cnt / kPwNS) & 1) ^ 1, 1, prof, pc[1]);
        const bool load_w = !p.pairs && (!p.w_resident || first);
        const int ws = p.w_resident ? kb : s;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], load_w ? 2 * kPwTile : kPwTile);
          tma_load_2d(a_st + s * kPwTile, &tm_a, &full[s], kb * 32, mt * 128);
          if (load_w) tma_load_2d(w_st + ws * kPwWStage, &tm_w, &full[s], kb * 32, nt * 128);
        }
        __syncwarp();
      }
      first = false;
    }
  } else if (wid == 21) {
    // ------------------------------------------------ MMA issuer
    uint32_t cnt = 0, use = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++use) {
      const int b = use % kPwNAcc;
      mbar_wait_prof(&acc_empty[b], ((use / kPwNAcc) & 1) ^ 1, 2, prof, pc[1]);
      tc_fence_after();
      const uint32_t d = tmem_base + b * 128;
      for (int kb = 0; kb < p.KB; ++kb, ++cnt) {
        const int s = cnt % kPwNS;
        mbar_wait_prof(&ready[s], (cnt / kPwNS) & 1, 3, prof, pc[2]);
        tc_fence_after();
        const int ws = p.w_resident ? kb : s;
        const uint32_t a_hi = tmem_base + kPwACol + s * 64;
        const uint32_t w_hi = sdesc_lo(smem_u32(w_st + ws * kPwWStage));
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            // the small cross terms first, hi*hi last
            const uint32_t a = a_hi + (pass == 0 ? 32 : 0);
            const uint32_t w = w_hi + (pass == 1 ? (kPwTile >> 4) : 0);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) mma_tf32_ts(d, a + 8 * k8, w + 2 * k8, kSdescHi, kIdescTf32, (kb | pass | k8) != 0);
          }
          tc_commit(&empty[s]);
          if (kb == p.KB - 1) tc_commit(&acc_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (wid < 4) {
    // ------------------------------------------------ A -> TMEM hi | lo (row = lane), W split in place
    uint32_t cnt = 0;
    bool first = true;
    const uint32_t a_t0 = tmem_base + (static_cast<uint32_t>(wid * 32) << 16) + kPwACol;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      for (int kb = 0; kb < p.KB; ++kb, ++cnt) {
        const int s = cnt % kPwNS;
        const int ws = p.w_resident ? kb : s;
        if (p.pairs && first) {
          // W' = diag(W, W): row n' = tid, 32 floats (8 cells of 16 B, swizzled by row & 7)
          unsigned char* row = w_st + ws * kPwWStage + tid * 128;
          const float* src = p.filter + (tid & 63) * 16;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((c >> 2) == (tid >> 6)) x = __ldg(reinterpret_cast<const float4*>(src) + (c & 3));
            float4 h, l;
            split4(x, &h, &l);
            float4* cell = reinterpret_cast<float4*>(row + ((c ^ (tid & 7)) << 4));
            *cell = h;
            *(cell + kPwTile / 16) = l;
          }
        }
        // (the stage's TMEM columns are free: the producer issued this load only after the MMAs
        // that last read them had retired)
        mbar_wait_prof(&full[s], (cnt / kPwNS) & 1, 4, prof, pc[1]);
        tc_fence_after();
        const long long ts0 = prof ? clock64() : 0;
        {
          const unsigned char* row = a_st + s * kPwTile + tid * 128;
          uint32_t h[32], l[32];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 x = *reinterpret_cast<const float4*>(row + ((k ^ (tid & 7)) << 4));
            float4 hh, ll;
            split4(x, &hh, &ll);
            h[4 * k] = __float_as_uint(hh.x); h[4 * k + 1] = __float_as_uint(hh.y);
            h[4 * k + 2] = __float_as_uint(hh.z); h[4 * k + 3] = __float_as_uint(hh.w);
            l[4 * k] = __float_as_uint(ll.x); l[4 * k + 1] = __float_as_uint(ll.y);
            l[4 * k + 2] = __float_as_uint(ll.z); l[4 * k + 3] = __float_as_uint(ll.w);
          }
          tmem_st32(a_t0 + s * 64, h);
          tmem_st32(a_t0 + s * 64 + 32, l);
        }
        const long long ts1 = prof ? clock64() : 0;
        if (!p.pairs && (!p.w_resident || first)) split_tile(w_st + ws * kPwWStage, tid);
        const long long ts2 = prof ? clock64() : 0;
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&ready[s]);
        if (prof) { pc[2] += ts1 - ts0; pc[3] += ts2 - ts1; pc[4] += clock64() - ts2; pc[5] += 1; }
      }
      first = false;
    }
  } else {
    // ------------------------------------------------ epilogue: warp -> TMEM lane quarter q, 32-column group g
    const int q = wid & 3, g = (wid - 4) >> 2;
    unsigned char* buf = epi + (wid - 4) * 4096;
    const int pw = p.N >> 5;   // packed words per row
    uint32_t use = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++use) {
      const int mt = it / p.n_tiles, nt = it - mt * p.n_tiles;
      const int b = use % kPwNAcc;
      const long long row = static_cast<long long>(mt) * 128 + q * 32 + lane;
      mbar_wait_prof(&acc_full[b], (use / kPwNAcc) & 1, 5, prof, pc[1]);
      tc_fence_after();
      const long long te0 = prof ? clock64() : 0;
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 128 + g * 32, v);
      // the accumulator is in registers: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
      const int n0 = nt * 128 + g * 32;
      uint32_t bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && n0 + 4 * k < p.N) bi = __ldg(reinterpret_cast<const float4*>(p.bias + ((n0 + 4 * k) & p.bias_mask)));
        float y0 = __fadd_rn(__uint_as_float(v[4 * k]), bi.x), y1 = __fadd_rn(__uint_as_float(v[4 * k + 1]), bi.y);
        float y2 = __fadd_rn(__uint_as_float(v[4 * k + 2]), bi.z), y3 = __fadd_rn(__uint_as_float(v[4 * k + 3]), bi.w);
        if (p.act == LCE_ACT_RELU) {
          y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
        } else if (p.act == LCE_ACT_RELU6) {
          y0 = fminf(fmaxf(y0, 0.f), 6.f); y1 = fminf(fmaxf(y1, 0.f), 6.f);
          y2 = fminf(fmaxf(y2, 0.f), 6.f); y3 = fminf(fmaxf(y3, 0.f), 6.f);
        } else if (p.act == LCE_ACT_RELU_N1_TO_1) {
          y0 = fminf(fmaxf(y0, -1.f), 1.f); y1 = fminf(fmaxf(y1, -1.f), 1.f);
          y2 = fminf(fmaxf(y2, -1.f), 1.f); y3 = fminf(fmaxf(y3, -1.f), 1.f);
        }
        v[4 * k] = __float_as_uint(y0); v[4 * k + 1] = __float_as_uint(y1);
        v[4 * k + 2] = __float_as_uint(y2); v[4 * k + 3] = __float_as_uint(y3);
        bits |= ((y0 < 0.0f ? 1u : 0u) | (y1 < 0.0f ? 2u : 0u) | (y2 < 0.0f ? 4u : 0u) | (y3 < 0.0f ? 8u : 0u)) << (4 * k);
      }
      const long long te1 = prof ? clock64() : 0;
      if (lane == 0) tma_store_wait_read();   // the previous item's store has read this buffer
      __syncwarp();
      const long long te2 = prof ? clock64() : 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<uint4*>(buf + lane * 128 + ((k ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm_out, buf, n0, mt * 128 + q * 32);
      if (p.packed != nullptr && row < p.M) p.packed[row * pw + nt * 4 + g] = static_cast<int32_t>(bits);
      if (prof) { pc[2] += te1 - te0; pc[3] += te2 - te1; pc[4] += clock64() - te2; pc[5] += 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  if (prof) {
    pc[0] = clock64() - prof_t0;
    for (int i = 0; i < 8; ++i) p.prof[wid * 8 + i] = pc[i];
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 20) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------ 7x7 / stride 2 stem on kind::tf32
// Bi-RealNet's first layer, CONV_2D 7x7 stride 2, 3 -> 64 channels: K = 147 (five 32-float K
// blocks, zero-padded), N = 64. As an FMA implicit GEMM it took 3.3 ms at batch 512, more than
// all the binary layers together. Same 3-pass split as above; what differs is the A operand: no
// im2col tensor exists, so no TMA -- thread = output pixel = TMEM lane gathers its own 147 input
// values (21 contiguous floats per filter row; neighbours share them through L1), splits them and
// writes hi | lo straight to TMEM. The filter (37 KB) is split into shared memory once per CTA.
//   warps 0-7   gather + split: two warps per TMEM lane quarter take alternate K blocks
//   warps 8-15  epilogue: (lane quarter, 32-column half)
//   warp 16     MMA issuer (+ TMEM allocation)
constexpr int kS7Threads = 544;
constexpr int kS7K = 147, kS7KB = 5, kS7N = 64;
constexpr int kS7WTile = 64 * 128;                 // 64 rows x 32 floats
constexpr int kS7NS = 3, kS7NAcc = 4;
constexpr size_t kS7Smem = kS7KB * 2 * kS7WTile + 8 * 4096 + 1024 + 1024;
struct Stem7Params {
  long long M;           // output pixels
  int H, W, OH, OW, ph, pw;
  int m_tiles, act;
  const float* in;       // [B][H][W][3]
  const float* filter;   // [64][7][7][3]
  const float* bias;
};
constexpr uint32_t kIdescTf32N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

template <int KBI>
__device__ __forceinline__ void stem7_gather(const float* base, int row_pitch, uint32_t rowmask, int e_lo, int e_hi,
                                             uint32_t (&h)[32], uint32_t (&l)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = KBI * 32 + j;
    float x = 0.0f;
    if (k < kS7K) {
      const int fy = k / 21, e = k - fy * 21;     // compile-time after unrolling
      if (((rowmask >> fy) & 1u) && e >= e_lo && e < e_hi) x = __ldg(base + fy * row_pitch + e);
    }
    const float hi = to_tf32(x);
    h[j] = __float_as_uint(hi);
    l[j] = __float_as_uint(to_tf32(x - hi));
  }
}

__global__ void __launch_bounds__(kS7Threads, 1)
stem7_tf32_kernel(const __grid_constant__ CUtensorMap tm_out, const Stem7Params p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* w_st = smem;                               // [kS7KB][hi | lo][64 x 128 B]
  unsigned char* epi = w_st + kS7KB * 2 * kS7WTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + 8 * 4096);
  uint64_t* ready = bars;                 // A stage written to TMEM   (128)
  uint64_t* empty = bars + kS7NS;         // MMAs of the stage retired (commit)
  uint64_t* acc_full = bars + 2 * kS7NS;
  uint64_t* acc_empty = acc_full + kS7NAcc;
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(acc_empty + kS7NAcc);

  const int tid = threadIdx.x;
  const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kS7NS; ++i) {
      mbar_init(&ready[i], 128);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kS7NAcc; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (wid == 16) tmem_alloc(tmem_base_s, 512);
  // the filter, split hi | lo, in the K-major SWIZZLE_128B image the MMA reads (K padded with zeros)
  for (int i = tid; i < kS7KB * 64 * 8; i += kS7Threads) {
    const int c = i & 7, row = (i >> 3) & 63, kb = i >> 9;
    float4 x;
    float* xp = reinterpret_cast<float*>(&x);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = kb * 32 + c * 4 + q;
      xp[q] = k < kS7K ? __ldg(p.filter + row * kS7K + k) : 0.0f;
    }
    float4 hh, ll;
    split4(x, &hh, &ll);
    float4* cell = reinterpret_cast<float4*>(w_st + kb * 2 * kS7WTile + row * 128 + ((c ^ (row & 7)) << 4));
    *cell = hh;
    *(cell + kS7WTile / 16) = ll;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  constexpr uint32_t kACol = kS7NAcc * kS7N;   // A stages behind the accumulators

  if (wid == 16) {
    // ------------------------------------------------ MMA issuer
    uint32_t cnt = 0, use = 0;
    for (int it = blockIdx.x; it < p.m_tiles; it += gridDim.x, ++use) {
      const int b = use % kS7NAcc;
      mbar_wait_tc(&acc_empty[b], ((use / kS7NAcc) & 1) ^ 1, 2);
      tc_fence_after();
      const uint32_t d = tmem_base + b * kS7N;
      for (int kb = 0; kb < kS7KB; ++kb, ++cnt) {
        const int s = cnt % kS7NS;
        mbar_wait_tc(&ready[s], (cnt / kS7NS) & 1, 3);
        tc_fence_after();
        const uint32_t a_hi = tmem_base + kACol + s * 64;
        const uint32_t w_hi = sdesc_lo(smem_u32(w_st + kb * 2 * kS7WTile));
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a = a_hi + (pass == 0 ? 32 : 0);
            const uint32_t w = w_hi + (pass == 1 ? (kS7WTile >> 4) : 0);
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) mma_tf32_ts(d, a + 8 * k8, w + 2 * k8, kSdescHi, kIdescTf32N64, (kb | pass | k8) != 0);
          }
          tc_commit(&empty[s]);
          if (kb == kS7KB - 1) tc_commit(&acc_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (wid < 8) {
    // ------------------------------------------------ gather + split: warp (quarter, parity) takes every other K block
    const int quarter = wid & 3, parity = wid >> 2;
    const uint32_t a_t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kACol;
    const int row_pitch = p.W * 3;
    const long long ohw = static_cast<long long>(p.OH) * p.OW;
    uint32_t cnt = 0;
    for (int it = blockIdx.x; it < p.m_tiles; it += gridDim.x) {
      const long long m = static_cast<long long>(it) * 128 + quarter * 32 + lane;
      const long long mm = m < p.M ? m : p.M - 1;          // rows past the end gather a valid pixel; never stored
      const int bi = static_cast<int>(mm / ohw);
      const int rem = static_cast<int>(mm - bi * ohw);
      const int oy = rem / p.OW, ox = rem - oy * p.OW;
      const int iy0 = oy * 2 - p.ph, ix0 = ox * 2 - p.pw;
      uint32_t rowmask = 0;
#pragma unroll
      for (int fy = 0; fy < 7; ++fy)
        if (static_cast<unsigned>(iy0 + fy) < static_cast<unsigned>(p.H)) rowmask |= 1u << fy;
      const int e_lo = 3 * max(0, -ix0), e_hi = 3 * min(7, p.W - ix0);
      const float* base = p.in + ((static_cast<long long>(bi) * p.H + iy0) * p.W + ix0) * 3;
      for (int kb = 0; kb < kS7KB; ++kb, ++cnt) {
        if ((cnt & 1u) != static_cast<uint32_t>(parity)) continue;
        const int s = cnt % kS7NS;
        uint32_t h[32], l[32];
        switch (kb) {
          case 0: stem7_gather<0>(base, row_pitch, rowmask, e_lo, e_hi, h, l); break;
          case 1: stem7_gather<1>(base, row_pitch, rowmask, e_lo, e_hi, h, l); break;
          case 2: stem7_gather<2>(base, row_pitch, rowmask, e_lo, e_hi, h, l); break;
          case 3: stem7_gather<3>(base, row_pitch, rowmask, e_lo, e_hi, h, l); break;
          default: stem7_gather<4>(base, row_pitch, rowmask, e_lo, e_hi, h, l); break;
        }
        mbar_wait_tc(&empty[s], ((cnt / kS7NS) & 1) ^ 1, 6);   // the MMAs that last read this TMEM stage retired
        tc_fence_after();
        tmem_st32(a_t0 + s * 64, h);
        tmem_st32(a_t0 + s * 64 + 32, l);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&ready[s]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: warp -> (TMEM lane quarter q, 32-column half g)
    const int q = wid & 3, g = (wid - 8) >> 2;
    unsigned char* buf = epi + (wid - 8) * 4096;
    uint32_t use = 0;
    for (int it = blockIdx.x; it < p.m_tiles; it += gridDim.x, ++use) {
      const int b = use % kS7NAcc;
      mbar_wait_tc(&acc_full[b], (use / kS7NAcc) & 1, 5);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * kS7N + g * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bi = __ldg(reinterpret_cast<const float4*>(p.bias + g * 32 + 4 * k));
        float y0 = __fadd_rn(__uint_as_float(v[4 * k]), bi.x), y1 = __fadd_rn(__uint_as_float(v[4 * k + 1]), bi.y);
        float y2 = __fadd_rn(__uint_as_float(v[4 * k + 2]), bi.z), y3 = __fadd_rn(__uint_as_float(v[4 * k + 3]), bi.w);
        if (p.act == LCE_ACT_RELU) {
          y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
        } else if (p.act == LCE_ACT_RELU6) {
          y0 = fminf(fmaxf(y0, 0.f), 6.f); y1 = fminf(fmaxf(y1, 0.f), 6.f);
          y2 = fminf(fmaxf(y2, 0.f), 6.f); y3 = fminf(fmaxf(y3, 0.f), 6.f);
        } else if (p.act == LCE_ACT_RELU_N1_TO_1) {
          y0 = fminf(fmaxf(y0, -1.f), 1.f); y1 = fminf(fmaxf(y1, -1.f), 1.f);
          y2 = fminf(fmaxf(y2, -1.f), 1.f); y3 = fminf(fmaxf(y3, -1.f), 1.f);
        }
        v[4 * k] = __float_as_uint(y0); v[4 * k + 1] = __float_as_uint(y1);
        v[4 * k + 2] = __float_as_uint(y2); v[4 * k + 3] = __float_as_uint(y3);
      }
      if (lane == 0) tma_store_wait_read();   // the previous item's store has read this buffer
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<uint4*>(buf + lane * 128 + ((k ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) tma_store_2d(&tm_out, buf, g * 32, it * 128 + q * 32);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 16) tmem_dealloc(tmem_base, 512);
}

}  // namespace pw
}  // namespace lce

#endif  // LCE_B200_PW_CUH_
