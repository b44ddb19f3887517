// lce_b200_imma.cuh -- the binary convolution's inner product on the int8 tensor pipe.
// Used for every plan with full 64-channel tiles and float / raw-accumulator output (the three
// model families' LceBconv2d layers, the BGEMM sweep); everything else, and every plan when
// LCE_B200_BCONV_IMMA=0, runs the XOR + POPC kernel of lce_b200_kernels.cuh that north_star
// describes. Both produce the reference's integers bit for bit (tests/test_gpu_parity.py runs
// the same cases through both); bench.py reports both.
//
// sm_100a has no b1 tensor instruction (ptxas lowers mma.sync ... b1 to masked int8 IMMAs that
// are no faster than the carry-save tree: profiles/r01_microbench_mma.jsonl), but the legacy
// int8 tensor pipe (mma.sync.m16n8k32, SASS IMMA.16832) sustains 2048 MAC/clk/SM, twice the
// ALU/XU ceiling of XOR+POPC. HBM, L2 and shared memory keep the reference's bitpacked words;
// only registers ever hold bytes. With activation bits a in {0,1} (u8) and weight bits mapped
// to w' = +1 / -1 (s8):
//     popc(a ^ w) = popc(w) + sum_k a_k * w'_k
// so the accumulators start at the per-channel popcount of the filter row and the tensor pipe
// adds the signed dot product -- the same integer the XOR+POPC kernel produces, bit for bit
// (out-of-bounds taps gather 0 words = "+1" padding exactly as before).
// To make the bits -> bytes expansion cheap, bit i of a nibble becomes the byte VALUE 2^i (one
// PRMT replicates the thread's byte into all four lanes of a register, one AND with 0x08040201
// isolates bit i in byte i) and the weight byte at that k position is +-(8 >> i): every product
// is +-8, the accumulators hold 8 * popc(a ^ w) exactly, and the epilogue's "<< 1" becomes ">> 2".
// Activations are expanded bits -> bytes in registers, per warp. The static weights are expanded
// once per plan into mma B-fragment order (8 bytes per lane, k-step and 8-channel sub-tile) and
// reach shared memory by TMA bulk copy, so a B fragment is one conflict-free 64-bit LDS.
//
// CTA = 128 pixels x 64 channels, 4 warps of 32 x 64 (2 x 8 m16n8 tiles, 64 accumulators per
// thread). Thread (gid = lane / 4, tig = lane % 4) expands byte `tig` of each 32-bit word: its
// low nibble feeds the k 0..15 half of the fragment, its high nibble the k 16..31 half -- any
// bit <-> k assignment is valid as long as A and B use the same one.
#ifndef LCE_B200_IMMA_CUH_
#define LCE_B200_IMMA_CUH_

#include "lce_b200_kernels.cuh"

namespace lce {

constexpr int kIBM = 128;
constexpr int kIThreads = 128;
// resident CTAs per SM the kernel is compiled for (register cap 65536 / (128 * n))
#ifdef LCE_IMMA_ALL4
__host__ __device__ constexpr int imma_ctas_per_sm(int) { return 4; }
#else
__host__ __device__ constexpr int imma_ctas_per_sm(int V) { return V == 4 ? 3 : 4; }
#endif
constexpr int kIBytesPerWord = kIBM * 4 + 8 * 32 * 8;  // smem per K word: packed A column + B fragments
// K is staged in one chunk when it fits, else through a TWO-deep ring: chunk c+1 is gathered /
// TMA-copied while chunk c is multiplied. The host sizes the chunks (imma_smem_budget):
// 3 CTAs per SM for the uint4 instances (153 registers), 4 for the narrower ones (128).

// byte `tig` of w -> two registers of four u8 each: bit i of the low / high nibble -> byte i with
// value 2^i (sel = tig * 0x1111 replicates the byte; 4 ALU instructions, no multiplies)
__device__ __forceinline__ void expand01(uint32_t w, uint32_t sel, uint32_t& lo, uint32_t& hi) {
  const uint32_t t = __byte_perm(w, 0u, sel);
  lo = t & 0x08040201u;
  hi = (t >> 4) & 0x08040201u;
}
__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// weight nibble -> four s8: byte i = +(8 >> i) for bit 0 (+1), -(8 >> i) for bit 1 (-1)
__host__ __device__ __forceinline__ uint32_t imma_weight_bytes(uint32_t nibble) {
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const int mag = 8 >> i;
    const int v = ((nibble >> i) & 1u) ? -mag : mag;
    r |= (static_cast<uint32_t>(v) & 0xFFu) << (8 * i);
  }
  return r;
}

// Plan-time: wtx[n_tile][k word][ns / 2][lane][ns & 1] = {b0, b1}, the B fragment (k-step = that
// word, channels ns*8 .. ns*8+7 of the tile) of `lane`; the fragments of two neighbouring
// sub-tiles sit side by side so one 128-bit LDS fetches both. Channels past the group's end read
// as 0 words.
__global__ void expand_weights_imma_kernel(const int32_t* __restrict__ filter,
                                           uint2* __restrict__ wtx, int cout_pg,
                                           int tiles_per_group, int taps, int Cw_pg,
                                           long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int ns = static_cast<int>(((idx >> 6) & 3) * 2 + (idx & 1));   // idx = ((kw*4 + ns/2)*32 + lane)*2 + (ns&1)
  const int lane = static_cast<int>((idx >> 1) & 31);
  const int Kw = taps * Cw_pg;
  const long long r = idx >> 8;
  const int kw = static_cast<int>(r % Kw);
  const int nt = static_cast<int>(r / Kw);
  const int g = nt / tiles_per_group, tg = nt - g * tiles_per_group;
  const int ch = tg * kBN + ns * 8 + (lane >> 2);
  uint32_t w = 0;
  if (ch < cout_pg)
    w = static_cast<uint32_t>(filter[(static_cast<long long>(g) * cout_pg + ch) * Kw + kw]);
  const uint32_t byte = (w >> (8 * (lane & 3))) & 0xFFu;
  wtx[idx] = make_uint2(imma_weight_bytes(byte & 0xFu), imma_weight_bytes(byte >> 4));
}

template <int V>
__device__ __forceinline__ void compute_chunk_imma(const typename VecT<V>::T* A_s,
                                                   const uint2* B_s, int nkv, int warp, int gid,
                                                   int tig, int lane, int (&acc)[2][8][4]) {
  const uint4* b4_base = reinterpret_cast<const uint4*>(B_s) + lane;
  using Vec = typename VecT<V>::T;
  const uint32_t sel = 0x1111u * static_cast<uint32_t>(tig);
  const Vec* a_base = A_s + warp * 32 + gid;
#pragma unroll 1
  for (int kv = 0; kv < nkv; ++kv) {
    uint32_t r[2][2][V];   // [m sub-tile][row gid / gid + 8][word]
#pragma unroll
    for (int ms = 0; ms < 2; ++ms) {
      load_words<V>(a_base + kv * kIBM + ms * 16, r[ms][0]);
      load_words<V>(a_base + kv * kIBM + ms * 16 + 8, r[ms][1]);
    }
    // two k-steps at a time: 16 live fragment registers instead of 8 * V
    constexpr int H = V < 2 ? V : 2;
#pragma unroll
    for (int qh = 0; qh < V; qh += H) {
      uint32_t a[H][2][4];
#pragma unroll
      for (int q = 0; q < H; ++q)
#pragma unroll
        for (int ms = 0; ms < 2; ++ms) {
          expand01(r[ms][0][qh + q], sel, a[q][ms][0], a[q][ms][2]);
          expand01(r[ms][1][qh + q], sel, a[q][ms][1], a[q][ms][3]);
        }
#pragma unroll
      for (int q = 0; q < H; ++q) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const uint4 b = b4_base[((kv * V + qh + q) * 4 + np) * 32];
          mma_u8s8(acc[0][2 * np], a[q][0], b.x, b.y);
          mma_u8s8(acc[1][2 * np], a[q][1], b.x, b.y);
          mma_u8s8(acc[0][2 * np + 1], a[q][0], b.z, b.w);
          mma_u8s8(acc[1][2 * np + 1], a[q][1], b.z, b.w);
        }
      }
    }
  }
}

// SAME padding with pad value 0 (reference.h:100-103), for this kernel's accumulator layout:
// an out-of-bounds tap contributes channels_in_per_group / 2 instead of popc(0 ^ w) -- times 8,
// the accumulators' scale.
__device__ __forceinline__ void zero_pad_correction_imma(const ConvKParams& p,
                                                         int (&acc)[2][8][4], long long m0,
                                                         int c_tile, int warp, int gid, int tig) {
  const int taps = p.KH * p.KW;
#pragma unroll
  for (int ms = 0; ms < 2; ++ms) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long m = m0 + warp * 32 + ms * 16 + h * 8 + gid;
      if (m >= p.M) continue;
      long long b;
      int oy, ox;
      split_pixel(p, m, &b, &oy, &ox);
      const int iy_lo = oy * p.sh - p.ph, ix_lo = ox * p.sw - p.pw;
      if (iy_lo >= 0 && ix_lo >= 0 && iy_lo + (p.KH - 1) * p.dh < p.H &&
          ix_lo + (p.KW - 1) * p.dw < p.W)
        continue;  // interior pixel
      for (int fy = 0; fy < p.KH; ++fy) {
        const int iy = iy_lo + fy * p.dh;
        const bool yin = static_cast<unsigned>(iy) < static_cast<unsigned>(p.H);
        for (int fx = 0; fx < p.KW; ++fx) {
          const int ix = ix_lo + fx * p.dw;
          if (yin && static_cast<unsigned>(ix) < static_cast<unsigned>(p.W)) continue;
          const int t = fy * p.KW + fx;
#pragma unroll
          for (int ns = 0; ns < 8; ++ns) {
            const int c = c_tile + ns * 8 + tig * 2;
            acc[ms][ns][h * 2] += 8 * (p.zp_half - p.tap_popc[static_cast<size_t>(c) * taps + t]);
            acc[ms][ns][h * 2 + 1] +=
                8 * (p.zp_half - p.tap_popc[static_cast<size_t>(c + 1) * taps + t]);
          }
        }
      }
    }
  }
}

// Accumulator (ms, ns, e) of thread (gid, tig): pixel m0 + warp*32 + ms*16 + (e/2)*8 + gid,
// channel c_tile + ns*8 + tig*2 + (e & 1). Full 64-channel tiles only (host-checked).
template <int OUT>
__device__ __forceinline__ void epilogue_imma(const ConvKParams& p, int (&acc)[2][8][4],
                                              long long m0, int c_tile, int warp, int gid,
                                              int tig) {
  const int cth = c_tile + tig * 2;
  float2 mul_r[8], bias_r[8];
  if (OUT == LCE_OUT_FLOAT) {
#pragma unroll
    for (int ns = 0; ns < 8; ++ns) {
      mul_r[ns] = *reinterpret_cast<const float2*>(p.mul + cth + ns * 8);
      bias_r[ns] = *reinterpret_cast<const float2*>(p.bias + cth + ns * 8);
    }
  }
  const bool has_res = OUT == LCE_OUT_FLOAT && p.residual != nullptr;
  const bool has_pk = OUT == LCE_OUT_FLOAT && p.packed_out != nullptr;
  // fused activation of the ADD as one clamp: y = min(max(y, lo), hi) (kernel_util.h:285-300)
  const int ract = p.residual_act;
  const float act_lo = ract == LCE_ACT_RELU_N1_TO_1 ? -1.0f : 0.0f;
  const float act_hi = ract == LCE_ACT_RELU ? __int_as_float(0x7f800000)
                                            : (ract == LCE_ACT_RELU6 ? 6.0f : 1.0f);
#pragma unroll
  for (int ms = 0; ms < 2; ++ms) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int lrow = warp * 32 + ms * 16 + h * 8 + gid;
      const long long m = m0 + lrow;
      const bool row_ok = m < p.M;
      const size_t e0 = static_cast<size_t>(m) * p.cout + cth;
      if (OUT == LCE_OUT_RAW_ACC) {
        if (row_ok) {
          int* o = static_cast<int*>(p.out) + e0;
#pragma unroll
          for (int ns = 0; ns < 8; ++ns)
            *reinterpret_cast<int2*>(o + ns * 8) =
                make_int2(acc[ms][ns][h * 2] >> 3, acc[ms][ns][h * 2 + 1] >> 3);
        }
        continue;
      }
      uint32_t bits0 = 0, bits1 = 0;
      if (row_ok) {
        float y[16];
#pragma unroll
        for (int ns = 0; ns < 8; ++ns) {
          // accumulators hold 8 * acc: OutputTransform's (acc << 1) is (acc8 >> 2), exactly
          y[2 * ns] = transform_float_x2(acc[ms][ns][h * 2] >> 2, p.clamp_min, p.clamp_max,
                                         mul_r[ns].x, bias_r[ns].x);
          y[2 * ns + 1] = transform_float_x2(acc[ms][ns][h * 2 + 1] >> 2, p.clamp_min, p.clamp_max,
                                            mul_r[ns].y, bias_r[ns].y);
        }
        if (has_res) {
          // (A TMA-staged copy of the shortcut tile was measured slower on every layer: the
          // extra 36 KB of shared memory cost a resident CTA. The rows are already in L2.)
          float2 rv[8];
          const float* r = p.residual + e0;
#pragma unroll
          for (int ns = 0; ns < 8; ++ns) rv[ns] = __ldg(reinterpret_cast<const float2*>(r + ns * 8));
#pragma unroll
          for (int ns = 0; ns < 8; ++ns) {
            y[2 * ns] = __fadd_rn(y[2 * ns], rv[ns].x);
            y[2 * ns + 1] = __fadd_rn(y[2 * ns + 1], rv[ns].y);
          }
          if (ract != LCE_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = fminf(fmaxf(y[j], act_lo), act_hi);
          }
        }
        float* o = static_cast<float*>(p.out) + e0;
#pragma unroll
        for (int ns = 0; ns < 8; ++ns)
          *reinterpret_cast<float2*>(o + ns * 8) = make_float2(y[2 * ns], y[2 * ns + 1]);
        if (has_pk) {
          // LceQuantize of the value just written: bit = value < 0 (bitpack.h:159)
#pragma unroll
          for (int ns = 0; ns < 4; ++ns) {
            bits0 |= ((y[2 * ns] < 0.0f ? 1u : 0u) | (y[2 * ns + 1] < 0.0f ? 2u : 0u)) << (ns * 8);
            bits1 |= ((y[2 * ns + 8] < 0.0f ? 1u : 0u) | (y[2 * ns + 9] < 0.0f ? 2u : 0u)) << (ns * 8);
          }
          bits0 <<= tig * 2;
          bits1 <<= tig * 2;
        }
      }
      if (has_pk) {  // uniform across the warp: every lane runs the shuffles
        bits0 |= __shfl_xor_sync(0xffffffffu, bits0, 1);
        bits1 |= __shfl_xor_sync(0xffffffffu, bits1, 1);
        bits0 |= __shfl_xor_sync(0xffffffffu, bits0, 2);
        bits1 |= __shfl_xor_sync(0xffffffffu, bits1, 2);
        if (row_ok && tig == 0) {
          int32_t* pk = p.packed_out + static_cast<size_t>(m) * p.cw_out + (c_tile >> 5);
          pk[0] = static_cast<int32_t>(bits0);
          pk[1] = static_cast<int32_t>(bits1);
        }
      }
    }
  }
}

template <int V, int OUT>
__global__ void __launch_bounds__(kIThreads, imma_ctas_per_sm(V))
bconv_imma_kernel(const ConvKParams p) {
  using Vec = typename VecT<V>::T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // two buffers of { A: [Kc_v][128] packed vectors, B: [Kc_v*V][8][32] int8 fragments }
  __shared__ __align__(8) uint64_t wbar[2];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int nt = blockIdx.y;
  const int g = static_cast<int>(fdiv(static_cast<uint32_t>(nt), p.fd_tpg));
  const int tg = nt - g * p.tiles_per_group;
  const long long m0 = static_cast<long long>(blockIdx.x) * kIBM;
  const int c_tile = g * p.cout_pg + tg * kBN;

  if (tid == 0) {
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // accumulators start at 8 * popc(filter row): acc8 = 8 * (popc(w) + sum a * w')
  int acc[2][8][4];
#pragma unroll
  for (int ns = 0; ns < 8; ++ns) {
    const int2 pw = *reinterpret_cast<const int2*>(p.wpop + c_tile + ns * 8 + tig * 2);
#pragma unroll
    for (int ms = 0; ms < 2; ++ms) {
      acc[ms][ns][0] = pw.x * 8; acc[ms][ns][1] = pw.y * 8;
      acc[ms][ns][2] = pw.x * 8; acc[ms][ns][3] = pw.y * 8;
    }
  }
  if (OUT == LCE_OUT_FLOAT && p.residual != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + warp * 32 + i * 8 + gid;
      if (m < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + m * p.cout + c_tile + tig * 16));
    }
  }

  const size_t chunk_bytes = static_cast<size_t>(p.Kc_v) * V * kIBytesPerWord;
  auto stage = [&](int ch) {
    unsigned char* base = smem_raw + (ch & 1) * chunk_bytes;
    Vec* A_b = reinterpret_cast<Vec*>(base);
    uint2* B_b = reinterpret_cast<uint2*>(A_b + static_cast<size_t>(p.Kc_v) * kIBM);
    const int kv0 = ch * p.Kc_v;
    const int kv1 = min(kv0 + p.Kc_v, p.Kv);
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(kv1 - kv0) * V * (8 * 32 * 8);
      mbar_arrive_expect_tx(&wbar[ch & 1], bytes);
      bulk_g2s(B_b, reinterpret_cast<const uint2*>(p.wt) +
                        (static_cast<size_t>(nt) * p.Kv + kv0) * V * (8 * 32),
               bytes, &wbar[ch & 1]);
    }
    gather_tile<V, kIBM, kIThreads>(p, A_b, m0, g, kv0, kv1, tid);
    cp_async_commit();
  };
  stage(0);
  for (int ch = 0; ch < p.n_chunks; ++ch) {
    const bool more = ch + 1 < p.n_chunks;
    if (more) stage(ch + 1);  // its buffer was released by the barrier that ended chunk ch - 1
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    mbar_wait(&wbar[ch & 1], (ch >> 1) & 1);
    __syncthreads();
    unsigned char* base = smem_raw + (ch & 1) * chunk_bytes;
    const Vec* A_b = reinterpret_cast<const Vec*>(base);
    const uint2* B_b = reinterpret_cast<const uint2*>(A_b + static_cast<size_t>(p.Kc_v) * kIBM);
    const int kv0 = ch * p.Kc_v;
    compute_chunk_imma<V>(A_b, B_b, min(kv0 + p.Kc_v, p.Kv) - kv0, warp, gid, tig, lane, acc);
    __syncthreads();
  }
  if (p.tap_popc != nullptr) zero_pad_correction_imma(p, acc, m0, c_tile, warp, gid, tig);
  epilogue_imma<OUT>(p, acc, m0, c_tile, warp, gid, tig);
}

}  // namespace lce
#endif
