// tc_probe.cu -- building-block probe for the tcgen05 binary-conv kernel (sm_100a).
// Nothing here is on the product path: it pins, on a real B200, the hardware conventions the
// kernel in compute_engine_b200/csrc/lce_b200_tc.cuh relies on, each against a CPU loop:
//   T1  tcgen05.mma kind::i8, A and B from shared memory (no-swizzle K-major core matrices):
//       instruction descriptor, shared-memory descriptor (LBO / SBO meaning), D layout in TMEM
//   T2  the same product with A written to TMEM by tcgen05.st.32x32b (row = lane, 4 k per column)
//   T3  TMA tensor maps: 1-D int32 map with an unaligned start and an out-of-bounds tail,
//       2-D float map with SWIZZLE_128B (load, raw dump, store)
//   T4  MMA issue rate for N = 64 / 128 / 256, A from shared memory and from TMEM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ inline uint64_t make_sdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout) << 61);
}
__host__ __device__ inline uint32_t make_idesc_i8(int M, int N, int a_signed, int b_signed) {
  return (2u << 4) | (static_cast<uint32_t>(a_signed) << 7) | (static_cast<uint32_t>(b_signed) << 10) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// no-swizzle K-major image: 8 x 16 B core matrices; along K adjacent (128 B), 8-row groups at (K/16)*128
__host__ __device__ inline int core_off(int row, int kb, int K) {
  return (row >> 3) * (K / 16) * 128 + (kb >> 4) * 128 + (row & 7) * 16 + (kb & 15);
}

// ------------------------------------------------------------------ T1 / T2
// mode 0: A from smem (SS).  mode 1: A from TMEM (TS), written with tcgen05.st.
// swap = 1 exchanges the LBO / SBO descriptor fields (hypothesis test).
__global__ void __launch_bounds__(128) probe_mma(const int8_t* __restrict__ A_img, const int8_t* __restrict__ B_img,
                                                 const int8_t* __restrict__ A_rows, int32_t* __restrict__ D, int mode,
                                                 int swap, int N, int K, int a_signed) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* As = smem;
  unsigned char* Bs = smem + 128 * K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * K / 16; i += 128) reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(A_img)[i];
  for (int i = tid; i < N * K / 16; i += 128) reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(B_img)[i];
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t d_t = tb;             // columns [0, N)
  const uint32_t a_t = tb + 256;       // columns [256, 256 + K/4)
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  if (mode == 1) {
    const int r = tid;
    for (int c0 = 0; c0 < K / 4; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t*>(A_rows + r * K + (c0 + j) * 4);
      tmem_st8(a_t + lane_base + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, N, a_signed, 1);
    const uint32_t lbo = 128, sbo = (K / 16) * 128;
    for (int ks = 0; ks < K / 32; ++ks) {
      const uint64_t bd = swap ? make_sdesc(smem_u32(Bs) + ks * 256, sbo, lbo, 0) : make_sdesc(smem_u32(Bs) + ks * 256, lbo, sbo, 0);
      if (mode == 0) {
        const uint64_t ad = swap ? make_sdesc(smem_u32(As) + ks * 256, sbo, lbo, 0) : make_sdesc(smem_u32(As) + ks * 256, lbo, sbo, 0);
        mma_i8_ss(d_t, ad, bd, idesc, ks > 0);
      } else {
        mma_i8_ts(d_t, a_t + ks * 8, bd, idesc, ks > 0);
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(d_t + lane_base + c0, v);
    for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = static_cast<int32_t>(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------ T4: issue rate
__global__ void __launch_bounds__(128) probe_rate(int mode, int N, int iters, long long* cycles, int alt) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01ff01ffu, 0, 0x02020202u);
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  {
    uint32_t v[8] = {0x01010101u, 0, 0x01000100u, 0, 1, 2, 3, 4};
    for (int c = 0; c < 32; c += 8) tmem_st8(tb + 256 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, N, 1, 1);
    const uint32_t As = smem_u32(smem), Bs = smem_u32(smem + 128 * 128);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int ks = it & 3;
      const uint64_t bd = make_sdesc(Bs + ks * 256, 128, 1024, 0);
      const uint32_t dd = tb + (alt ? (it % alt) * N : 0);   // alt independent accumulators (alt * N <= 256)
      if (mode == 0) mma_i8_ss(dd, make_sdesc(As + ks * 256, 128, 1024, 0), bd, idesc, 1);
      else mma_i8_ts(dd, tb + 256 + ks * 8, bd, idesc, 1);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------ T5: issue interval vs accumulator
// 16 MMAs unrolled per loop trip, descriptors precomputed (no scalar work between issues); ALT
// accumulators taken in turn: if the ~100-cycle interval of T4 is the latency of the dependent
// accumulate (same D back to back), it must shrink with ALT >= 2.
template <int ALT>
__global__ void __launch_bounds__(128) probe_rate2(int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01ff01ffu, 0, 0x02020202u);
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  {
    uint32_t v[8] = {0x01010101u, 0, 0x01000100u, 0, 1, 2, 3, 4};
    for (int c = 0; c < 64; c += 8) tmem_st8(tb + 256 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0;
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, N, 1, 1);
    const uint32_t Bs = smem_u32(smem + 128 * 128);
    const uint64_t bd0 = make_sdesc(Bs, 128, 1024, 0);
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    t0 = clock64();
    if (pred) {
      for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          mma_i8_ts(tb + (j % ALT) * N, tb + 256 + (j & 7) * 8, bd0 + (j & 3) * 16, idesc, 1);
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// ------------------------------------------------------------------ T8: MMA rate under TMEM / shared-memory traffic
// The real kernels never reach T5's N/2 cycles per MMA (100-150 measured). Which neighbour slows
// the tensor pipe down? One thread issues back-to-back TS MMAs (N = 128, two accumulators in
// turn) while 8 other warps keep doing, until told to stop:
//   bit 0: tcgen05.st x32 into TMEM columns the MMAs do not touch (what the expanders do)
//   bit 1: tcgen05.ld x32 of other TMEM columns                    (what the epilogue does)
//   bit 2: 128-bit shared-memory stores + loads                     (staging traffic)
//   bit 3: a dense integer ALU stream                                (expander / epilogue arithmetic)
// `real` = 1: the issuing thread derives every descriptor from loop-carried values like the kernels do.
__global__ void __launch_bounds__(288) probe_contend(int mode, int iters, long long* cycles, int real, int no_mma = 0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (128 + 256) * 128 / 16; i += 288) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01ff01ffu, 0, 0x02020202u);
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp < 4) {
    uint32_t v[8] = {0x01010101u, 0, 0x01000100u, 0, 1, 2, 3, 4};
    for (int c = 0; c < 64; c += 8) tmem_st8(tb + 256 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0;
  if (warp == 8) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, 128, 1, 1);
    const uint32_t Bs = smem_u32(smem + 128 * 128);
    const uint64_t bd0 = make_sdesc(Bs, 128, 1024, 0);
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    t0 = clock64();
    if (no_mma) {                       // same duration, tensor pipe idle: the neighbours' baseline
      while (clock64() - t0 < 64LL * iters) {}
      if (lane == 0) { cycles[blockIdx.x] = clock64() - t0; stop = 1; }
    } else if (pred) {
      if (!real) {
        for (int it = 0; it < iters; it += 16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mma_i8_ts(tb + (j & 1) * 128, tb + 256 + (j & 7) * 8, bd0 + (j & 3) * 16, idesc, 1);
        }
      } else {
        uint32_t as = real - 1, d = 0;          // opaque to the compiler: stage / accumulator indices
        for (int it = 0; it < iters; it += 8) {
          const uint32_t a0 = tb + 256 + (as & 3) * 64 * 0 + (as & 1) * 0;
          const uint64_t b0 = bd0 + static_cast<uint64_t>((as & 3) * 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) mma_i8_ts(tb + (d & 1) * 128, a0 + j * 8, b0 + (j & 3) * 16, idesc, (it | j) != 0);
          as += 1;
          if ((as & 7) == 0) d ^= 1;
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    if (!no_mma) {
      mbar_wait(&bar, 0);
      if (lane == 0) { cycles[blockIdx.x] = clock64() - t0; stop = 1; }
    }
  } else {
    // warps 0-7: lane quarter = warp & 3; columns 320..511 are free for them
    const uint32_t tq = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t v[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    uint4* sp = reinterpret_cast<uint4*>(smem + (128 + 256) * 128) + tid;   // 8 KB scratch behind the operands
    uint32_t acc = 0;
    long long loops = 0;
    while (!stop) {
      ++loops;
      if (mode & 1) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) tmem_st8(tq + 320 + (warp >> 2) * 32 + c, v);
        tmem_st_wait();
      }
      if (mode & 2) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) { uint32_t r[8]; tmem_ld8(tq + 384 + (warp >> 2) * 32 + c, r); acc += r[0]; }
      }
      if (mode & 4) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sp[k * 256] = make_uint4(acc, k, 0, 0); acc += sp[((k + 1) & 7) * 256].x; }
      }
      if (mode & 8) {
#pragma unroll
        for (int k = 0; k < 64; ++k) acc = acc * 0x9E3779B1u + (acc >> 7) + k;
      }
      if (mode == 0) __nanosleep(200);
    }
    if (acc == 0x12345678u) cycles[blockIdx.x] = 0;
    if (tid == 0) cycles[148 + blockIdx.x] = loops;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------ T6: per-stage overhead of the issuing thread
// One "stage" = [optional: two try_waits on already-completed mbarriers] + fence + elect + 8 MMAs
// with descriptors derived from loop-carried values (as the real kernel does) + [optional: two
// tcgen05.commit]. Cycles per stage against the 8 * N / 2 the tensor pipe needs.
__global__ void __launch_bounds__(128) probe_stage(int N, int stages, int do_wait, int do_commit, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar, done_bar[2], dummy[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x01010101u, 0x01ff01ffu, 0, 0x02020202u);
  fence_proxy_async();
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done_bar[0], 1); mbar_init(&done_bar[1], 1);
    mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0);
  {
    uint32_t v[8] = {0x01010101u, 0, 0x01000100u, 0, 1, 2, 3, 4};
    for (int c = 0; c < 256; c += 8) tmem_st8(tb + 256 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0;
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_i8(128, N, 1, 1);
    const uint32_t Bs = smem_u32(smem + 128 * 128);
    uint32_t as = 0;
    t0 = clock64();
    for (int st = 0; st < stages; ++st) {
      if (do_wait) {               // parity 1 of a fresh barrier: complete at once
        mbar_wait(&done_bar[0], 1);
        mbar_wait(&done_bar[1], 1);
      }
      tc_fence_after();
      const uint64_t bd0 = make_sdesc(Bs + (as & 1) * 2048, 128, 1024, 0);
      const uint32_t a0 = tb + 256 + as * 64;
      uint32_t pred;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_i8_ts(tb, a0 + j * 8, bd0 + (j & 3) * 16, idesc, 1);
        if (do_commit) {
          tc_commit(&dummy[0]);
          tc_commit(&dummy[1]);
        }
      }
      __syncwarp();
      if (++as == 4) as = 0;
    }
        uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    if (pred) tc_commit(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------ T7: kind::tf32 on TMA-loaded SWIZZLE_128B operands
// D[128][N] = A[128][K] * B[N][K]^T in fp32 inputs (K-major rows of 32 floats = 128 B per swizzle
// row), as the fp32 1x1 convolutions around the binary path would use it: does the tensor core
// truncate or round fp32 -> tf32, and how accurate is the 3-pass hi/lo split?
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) probe_tf32(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                                                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                                                  float* D, int N, int K, int passes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar, lbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nkb = K / 32;                       // 128-byte K blocks
  unsigned char* As = smem;                     // [nkb][128 rows][128 B]
  unsigned char* Al = As + nkb * 16384;
  unsigned char* Bs = Al + nkb * 16384;         // [nkb][N rows][128 B]
  unsigned char* Bl = Bs + nkb * N * 128;
  if (tid == 0) {
    mbar_init(&bar, 1); mbar_init(&lbar, 1);
    fence_barrier_init();
    mbar_expect_tx(&lbar, 2 * nkb * (16384 + N * 128));
    for (int kb = 0; kb < nkb; ++kb) {
      const CUtensorMap* maps[4] = {&tmA, &tmAlo, &tmB, &tmBlo};
      unsigned char* dst[4] = {As + kb * 16384, Al + kb * 16384, Bs + kb * N * 128, Bl + kb * N * 128};
      for (int q = 0; q < 4; ++q)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst[q])),
                     "l"(reinterpret_cast<uint64_t>(maps[q])), "r"(smem_u32(&lbar)), "r"(kb * 32), "r"(0) : "memory");
    }
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  mbar_wait(&lbar, 0);
  if (tid == 0) {
    tc_fence_after();
    // c_format F32 (1), a/b format TF32 (2), K-major, N, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    uint32_t acc = 0;
    for (int pass = 0; pass < passes; ++pass) {        // hi*hi, hi*lo, lo*hi
      const unsigned char* a = pass == 2 ? Al : As;
      const unsigned char* b = pass == 1 ? Bl : Bs;
      for (int kb = 0; kb < nkb; ++kb)
        for (int k8 = 0; k8 < 4; ++k8) {
          const uint64_t ad = make_sdesc(smem_u32(a + kb * 16384) + k8 * 32, 16, 1024, 2);
          const uint64_t bd = make_sdesc(smem_u32(b + kb * N * 128) + k8 * 32, 16, 1024, 2);
          mma_tf32_ss(tb, ad, bd, idesc, acc);
          acc = 1;
        }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tb + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// ------------------------------------------------------------------ T9: kind::tf32 issue interval
// Back-to-back tf32 MMAs (M 128, N 128, K 8), A from shared memory (SS) or from TMEM (TS), two
// accumulators in turn. Nominal: half the bf16 rate = 64 cycles.
__device__ __forceinline__ void mma_tf32_ts_probe(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) probe_tf32_rate(int ts, int N, int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3f800000u, 0x3f000000u, 0, 0x40000000u);
  fence_proxy_async();
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  {
    uint32_t v[8] = {0x3f800000u, 0, 0x3f000000u, 0, 0x40000000u, 0, 0, 0x3f800000u};
    for (int c = 0; c < 64; c += 8) tmem_st8(tb + 256 + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  long long t0 = 0;
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    const uint64_t ad0 = make_sdesc(smem_u32(smem), 128, 1024, 0);
    const uint64_t bd0 = make_sdesc(smem_u32(smem + 128 * 128), 128, 1024, 0);
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    t0 = clock64();
    if (pred) {
      for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (ts) mma_tf32_ts_probe(tb + (j & 1) * N, tb + 256 + (j & 7) * 8, bd0 + (j & 3) * 16, idesc, 1);
          else mma_tf32_ss(tb + (j & 1) * N, ad0 + (j & 3) * 16, bd0 + (j & 3) * 16, idesc, 1);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------ T3: TMA
__global__ void probe_tma1d(const __grid_constant__ CUtensorMap tm, int c0, int nbox, int32_t* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    fence_proxy_async();
    mbar_expect_tx(&bar, nbox * 1024);
    for (int i = 0; i < nbox; ++i)
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(smem_u32(smem + i * 1024)),
                   "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(c0 + i * 256) : "memory");
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < nbox * 256; i += blockDim.x) out[i] = reinterpret_cast<int32_t*>(smem)[i];
}

__global__ void probe_tma2d(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, int col0, int row0,
                            float* raw_dump) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  float* in_s = reinterpret_cast<float*>(smem);           // 32 rows x 128 B
  float* out_s = reinterpret_cast<float*>(smem + 4096);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    mbar_expect_tx(&bar, 4096);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(in_s)),
                 "l"(reinterpret_cast<uint64_t>(&tm_in)), "r"(smem_u32(&bar)), "r"(col0), "r"(row0) : "memory");
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) raw_dump[i] = in_s[i];
  // thread t = row t: read its row through the hypothesised swizzle, write 2x into out_s the same way
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    for (int ch = 0; ch < 8; ++ch) {
      const int phys = r * 128 + ((ch ^ (r & 7)) << 4);
      float4 v = *reinterpret_cast<const float4*>(smem + phys);
      v.x *= 2.f; v.y *= 2.f; v.z *= 2.f; v.w *= 2.f;
      *reinterpret_cast<float4*>(smem + 4096 + phys) = v;
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm_out)),
                 "r"(smem_u32(out_s)), "r"(col0), "r"(row0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("cuTensorMapEncodeTiled not found\n"); exit(2); }
  return reinterpret_cast<EncodeTiledFn>(fn);
}

static int run_mma_case(int mode, int swap, int N, int K, int a_signed, bool verbose) {
  std::vector<int8_t> A(128 * K), B(N * K), Aimg(128 * K), Bimg(N * K);
  srand(1234 + N + K);
  for (auto& v : A) v = a_signed ? static_cast<int8_t>((rand() % 17) - 8) : static_cast<int8_t>(rand() % 9);
  for (auto& v : B) v = static_cast<int8_t>((rand() % 17) - 8);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) Aimg[core_off(r, k, K)] = A[r * K + k];
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bimg[core_off(n, k, K)] = B[n * K + k];
  int8_t *dA, *dB, *dAr; int32_t* dD;
  CK(cudaMalloc(&dA, Aimg.size())); CK(cudaMalloc(&dB, Bimg.size())); CK(cudaMalloc(&dAr, A.size()));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAr, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xEE, 128 * N * 4));
  const size_t smem = (128 + N) * K;
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  probe_mma<<<1, 128, smem>>>(dA, dB, dAr, dD, mode, swap, N, K, a_signed);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<int32_t> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += static_cast<int>(A[r * K + k]) * static_cast<int>(B[n * K + k]);
      if (ref != D[r * N + n]) {
        if (bad < 6 && verbose) printf("    D[%d][%d] = %d, want %d\n", r, n, D[r * N + n], ref);
        ++bad;
      }
    }
  printf("  mma mode=%s swap=%d N=%d K=%d a_signed=%d : %d / %d mismatches\n", mode ? "TS" : "SS", swap, N, K, a_signed, bad, 128 * N);
  cudaFree(dA); cudaFree(dB); cudaFree(dAr); cudaFree(dD);
  return bad;
}

// layout discovery when a hypothesis fails: B = "identity" rows so that D[r][n] = A[r][k = n]
static void discover_ts() {
  const int N = 64, K = 32;
  std::vector<int8_t> A(128 * K), Bimg(N * K, 0), Aimg(128 * K, 0);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) A[r * K + k] = static_cast<int8_t>(1 + k + 32 * (r % 3));
  for (int n = 0; n < K; ++n) Bimg[core_off(n, n, K)] = 1;
  int8_t *dA, *dB, *dAr; int32_t* dD;
  CK(cudaMalloc(&dA, Aimg.size())); CK(cudaMalloc(&dB, Bimg.size())); CK(cudaMalloc(&dAr, A.size())); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAr, A.data(), A.size(), cudaMemcpyHostToDevice));
  probe_mma<<<1, 128, (128 + N) * K>>>(dA, dB, dAr, dD, 1, 0, N, K, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  for (int r : {0, 1, 2, 33, 127}) {
    printf("  discover TS row %3d (want 1+k+32*(r%%3)):", r);
    for (int n = 0; n < 32; ++n) printf(" %d", D[r * N + n]);
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dAr); cudaFree(dD);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  printf("T1: SS mode\n");
  int ss0 = run_mma_case(0, 0, 64, 128, 1, true);
  int ss1 = run_mma_case(0, 1, 64, 128, 1, false);
  printf("  => LBO/SBO hypothesis: %s\n", ss0 == 0 ? "as written (LBO = K-adjacent core matrices, SBO = 8-row groups)" : (ss1 == 0 ? "SWAPPED" : "NEITHER"));
  const int swap = (ss0 != 0 && ss1 == 0) ? 1 : 0;
  run_mma_case(0, swap, 128, 128, 1, true);
  run_mma_case(0, swap, 256, 64, 1, true);
  run_mma_case(0, swap, 64, 128, 0, true);   // unsigned A
  run_mma_case(0, swap, 16, 32, 1, true);
  printf("T2: TS mode (A in TMEM)\n");
  int ts = run_mma_case(1, swap, 64, 128, 1, true);
  run_mma_case(1, swap, 128, 128, 1, true);
  run_mma_case(1, swap, 256, 32, 0, true);
  if (ts != 0) discover_ts();

  printf("T3: TMA\n");
  EncodeTiledFn encode = get_encode();
  {
    const int n = 256 * 5 + 77;
    std::vector<int32_t> h(n);
    for (int i = 0; i < n; ++i) h[i] = i * 7 + 1;
    int32_t *d, *o;
    CK(cudaMalloc(&d, ((n * 4 + 15) / 16) * 16)); CK(cudaMalloc(&o, 1024 * 4));
    CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t gdim[1] = {static_cast<cuuint64_t>(n)};
    cuuint64_t gstr[1] = {0};
    cuuint32_t box[1] = {256};
    cuuint32_t es[1] = {1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 1, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode 1d: %d\n", static_cast<int>(r));
    for (int c0 : {0, 36, 256 * 4 + 4, -8}) {
      probe_tma1d<<<1, 128, 4 * 1024>>>(tm, c0, 2, o);
      CK(cudaDeviceSynchronize());
      std::vector<int32_t> got(512);
      CK(cudaMemcpy(got.data(), o, 512 * 4, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int i = 0; i < 512; ++i) {
        const int g = c0 + i;
        const int32_t want = (g >= 0 && g < n) ? h[g] : 0;
        if (got[i] != want) { if (bad < 4) printf("    1d c0=%d i=%d got %d want %d\n", c0, i, got[i], want); ++bad; }
      }
      printf("  tma 1d c0=%d: %d mismatches\n", c0, bad);
    }
    cudaFree(d); cudaFree(o);
  }
  {
    const int R = 70, C = 64;   // 70 rows: the last box is partly out of bounds
    std::vector<float> h(R * C);
    for (int i = 0; i < R * C; ++i) h[i] = static_cast<float>(i);
    float *d, *o, *raw;
    CK(cudaMalloc(&d, R * C * 4)); CK(cudaMalloc(&o, R * C * 4)); CK(cudaMalloc(&raw, 4096));
    CK(cudaMemcpy(d, h.data(), R * C * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(o, 0, R * C * 4));
    CUtensorMap tmi, tmo;
    cuuint64_t gdim[2] = {C, R};
    cuuint64_t gstr[1] = {C * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r1 = encode(&tmi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&tmo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, o, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode 2d: %d %d\n", static_cast<int>(r1), static_cast<int>(r2));
    for (int row0 : {0, 32, 64}) {
      const int col0 = 32;
      probe_tma2d<<<1, 128, 8192>>>(tmi, tmo, col0, row0, raw);
      CK(cudaDeviceSynchronize());
      std::vector<float> rw(1024), out(R * C);
      CK(cudaMemcpy(rw.data(), raw, 4096, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out.data(), o, R * C * 4, cudaMemcpyDeviceToHost));
      int bad_sw = 0, bad_st = 0;
      for (int r = 0; r < 32; ++r)
        for (int c = 0; c < 32; ++c) {
          const int phys = r * 32 + (((c >> 2) ^ (r & 7)) << 2) + (c & 3);
          const float want = (row0 + r < R) ? h[(row0 + r) * C + col0 + c] : 0.f;
          if (rw[phys] != want) { if (bad_sw < 4) printf("    swizzle r=%d c=%d got %g want %g\n", r, c, rw[phys], want); ++bad_sw; }
          if (row0 + r < R && out[(row0 + r) * C + col0 + c] != 2.f * want) ++bad_st;
        }
      printf("  tma 2d row0=%d: swizzle-hypothesis mismatches %d, store mismatches %d\n", row0, bad_sw, bad_st);
    }
    cudaFree(d); cudaFree(o); cudaFree(raw);
  }

  printf("T4: MMA issue rate (148 CTAs, 1 per SM)\n");
  {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int alt : {0, 2, 4})
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {32, 64, 128, 256}) {
        if (alt * N > 256) continue;
        const int iters = 8192;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe_rate<<<148, 128, (128 + 256) * 128>>>(mode, N, 64, dc, alt);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        probe_rate<<<148, 128, (128 + 256) * 128>>>(mode, N, iters, dc, alt);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> c(148);
        CK(cudaMemcpy(c.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
        const double macs = 148.0 * iters * 128.0 * N * 32.0;
        printf("  alt=%d mode=%s N=%3d: %.3f ms, %.1f cycles/MMA (SM0), %.0f MAC/clk/SM, %.1f int8 TOP/s\n", alt, mode ? "TS" : "SS", N, ms,
               static_cast<double>(c[0]) / iters, 128.0 * N * 32.0 * iters / static_cast<double>(c[0]), 2.0 * macs / (ms * 1e-3) / 1e12);
      }
    cudaFree(dc);
  }
  printf("T5: issue interval, unrolled x16, ALT accumulators in turn (TS mode)\n");
  {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    auto run = [&](auto kern, int alt, int N) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      const int iters = 8192;
      kern<<<148, 128, (128 + 256) * 128>>>(N, 64, dc);
      CK(cudaDeviceSynchronize());
      kern<<<148, 128, (128 + 256) * 128>>>(N, iters, dc);
      CK(cudaDeviceSynchronize());
      std::vector<long long> c(148);
      CK(cudaMemcpy(c.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
      printf("  alt=%d N=%3d: %.1f cycles/MMA, %.0f MAC/clk/SM\n", alt, N, static_cast<double>(c[0]) / iters,
             128.0 * N * 32.0 * iters / static_cast<double>(c[0]));
    };
    for (int N : {32, 64, 128}) {
      run(probe_rate2<1>, 1, N);
      run(probe_rate2<2>, 2, N);
      if (N <= 64) run(probe_rate2<4>, 4, N);
    }
    run(probe_rate2<1>, 1, 256);
    cudaFree(dc);
  }
  printf("T6: cycles per 8-MMA stage of the issuing thread (tensor time = 4 * N)\n");
  {
    long long* dc;
    CK(cudaMalloc(&dc, 148 * 8));
    CK(cudaFuncSetAttribute(probe_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int N : {64, 128})
      for (int w = 0; w < 2; ++w)
        for (int c = 0; c < 2; ++c) {
          const int stages = 1024;
          probe_stage<<<148, 128, (128 + 256) * 128>>>(N, 16, w, c, dc);
          CK(cudaDeviceSynchronize());
          probe_stage<<<148, 128, (128 + 256) * 128>>>(N, stages, w, c, dc);
          CK(cudaDeviceSynchronize());
          std::vector<long long> cy(148);
          CK(cudaMemcpy(cy.data(), dc, 148 * 8, cudaMemcpyDeviceToHost));
          printf("  N=%3d waits=%d commits=%d: %.0f cycles/stage (tensor %d)\n", N, w, c, static_cast<double>(cy[0]) / stages, 4 * N);
        }
    cudaFree(dc);
  }
  printf("T8: cycles per TS MMA (N = 128) while 8 other warps generate traffic\n");
  {
    CK(cudaFuncSetAttribute(probe_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long* dc;
    CK(cudaMalloc(&dc, 2 * 148 * sizeof(long long)));
    const char* names[8] = {"idle neighbours", "tcgen05.st", "tcgen05.ld", "tcgen05.st + ld", "shared ld/st", "st + shared", "ld + shared", "st + ld + shared"};
    for (int real = 0; real < 2; ++real)
      for (int mode : {0, 1, 2, 3, 4, 7, 8, 15}) {
        const int iters = 4096;
        probe_contend<<<148, 288, (128 + 256) * 128 + 9 * 1024 * 4>>>(mode, 64, dc, real);
        CK(cudaDeviceSynchronize());
        probe_contend<<<148, 288, (128 + 256) * 128 + 9 * 1024 * 4>>>(mode, iters, dc, real);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost));
        printf("  %s %-18s%s %.1f cycles/MMA (SM0), %.0f MAC/clk/SM\n", real ? "runtime descriptors," : "constant descriptors,",
               names[mode & 7], (mode & 8) ? " + ALU stream" : "", static_cast<double>(h[0]) / iters, 128.0 * 128 * 32 * iters / h[0]);
      }
    CK(cudaFree(dc));
  }
  printf("T10: do running MMAs slow the neighbours' TMEM traffic? (warp 0: loops of 4 x tcgen05.st/ld x8 + wait)\n");
  {
    long long* dc;
    CK(cudaMalloc(&dc, 2 * 148 * sizeof(long long)));
    for (int mode : {1, 2, 4}) {
      double rate[2];
      for (int no_mma = 0; no_mma < 2; ++no_mma) {
        const int iters = 8192;
        probe_contend<<<148, 288, (128 + 256) * 128 + 9 * 1024 * 4>>>(mode, iters, dc, 0, no_mma);
        CK(cudaDeviceSynchronize());
        long long h[296];
        CK(cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost));
        rate[no_mma] = static_cast<double>(h[0]) / h[148];
      }
      printf("  %-12s %.0f cycles per loop while MMAs run back to back, %.0f with the tensor pipe idle\n",
             mode == 1 ? "tcgen05.st" : mode == 2 ? "tcgen05.ld" : "shared ld/st", rate[0], rate[1]);
    }
    CK(cudaFree(dc));
  }
  printf("T9: kind::tf32 issue interval (M 128, K 8)\n");
  {
    CK(cudaFuncSetAttribute(probe_tf32_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long* dc;
    CK(cudaMalloc(&dc, 148 * sizeof(long long)));
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {64, 128, 256}) {
        const int iters = 4096;
        probe_tf32_rate<<<148, 128, (128 + 256) * 128>>>(ts, N, 64, dc);
        CK(cudaDeviceSynchronize());
        probe_tf32_rate<<<148, 128, (128 + 256) * 128>>>(ts, N, iters, dc);
        CK(cudaDeviceSynchronize());
        long long h[148];
        CK(cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost));
        printf("  %s N=%3d: %.1f cycles/MMA (SM0), %.0f MAC/clk/SM\n", ts ? "TS" : "SS", N, static_cast<double>(h[0]) / iters,
               128.0 * N * 8 * iters / h[0]);
      }
    CK(cudaFree(dc));
  }
  printf("T7: kind::tf32, TMA SWIZZLE_128B operands\n");
  {
    const int N = 128, K = 64;
    std::vector<float> A(128 * K), B(N * K), Ah(128 * K), Al(128 * K), Bh(N * K), Bl(N * K);
    srand(7);
    auto rnd = [] { return (static_cast<float>(rand()) / RAND_MAX - 0.5f) * 4.0f; };
    auto trunc_tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    auto rn_tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    for (size_t i = 0; i < A.size(); ++i) { A[i] = rnd(); Ah[i] = trunc_tf32(A[i]); Al[i] = A[i] - Ah[i]; }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = rnd(); Bh[i] = trunc_tf32(B[i]); Bl[i] = B[i] - Bh[i]; }
    float *dA, *dAl, *dB, *dBl, *dAh, *dBh, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4)); CK(cudaMalloc(&dAh, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dBl, B.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 4));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh.data(), B.size() * 4, cudaMemcpyHostToDevice));
    auto mk = [&](float* ptr, int rows) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
      cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K) * 4};
      cuuint32_t box[2] = {32, static_cast<cuuint32_t>(rows)};
      cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) printf("  encode failed %d\n", static_cast<int>(r));
      return tm;
    };
    CK(cudaFuncSetAttribute(probe_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const size_t smem = 2 * (K / 32) * (16384 + N * 128);
    std::vector<float> D(128 * N);
    auto report = [&](const char* what) {
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double e_full = 0, e_tr = 0, e_rn = 0, mag = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          double full = 0, tr = 0, rn = 0;
          for (int k = 0; k < K; ++k) {
            full += static_cast<double>(A[r * K + k]) * B[n * K + k];
            tr += static_cast<double>(trunc_tf32(A[r * K + k])) * trunc_tf32(B[n * K + k]);
            rn += static_cast<double>(rn_tf32(A[r * K + k])) * rn_tf32(B[n * K + k]);
          }
          const double d = D[r * N + n];
          e_full = std::max(e_full, std::abs(d - full)); e_tr = std::max(e_tr, std::abs(d - tr)); e_rn = std::max(e_rn, std::abs(d - rn));
          mag = std::max(mag, std::abs(full));
        }
      printf("  %s: max |D - exact| = %.3g, |D - truncated-inputs| = %.3g, |D - rounded-inputs| = %.3g (max |D| %.3g)\n", what, e_full,
             e_tr, e_rn, mag);
    };
    probe_tf32<<<1, 128, smem>>>(mk(dA, 128), mk(dAl, 128), mk(dB, N), mk(dBl, N), dD, N, K, 1);
    report("1 pass, raw fp32 operands   ");
    probe_tf32<<<1, 128, smem>>>(mk(dAh, 128), mk(dAl, 128), mk(dBh, N), mk(dBl, N), dD, N, K, 3);
    report("3 passes, hi/lo split        ");
  }
  printf("probe done\n");
  return 0;
}
