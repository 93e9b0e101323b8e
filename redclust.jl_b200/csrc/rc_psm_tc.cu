// rc_psm_tc.cu -- co-clustering counts of the posterior similarity matrix on the 5th-generation tensor cores.
//   PSM = sum(adjacencymatrix.(clusts)) ./ numsamples     /root/reference/src/mcmc.jl:560, src/utils.jl:59-63
// counts = Z * Z' where Z is the n x (S * SW) one-hot image of the label matrix (SW = slots per sample, the
// smallest of 32 / 64 / 128 that holds the largest label).  Z is never materialised in HBM: every CTA owns a
// 128 x 256 tile of counts and keeps two stages of a 128-row and a 256-row slice of Z in shared memory, in the
// K-major no-swizzle canonical layout of the UMMA shared-memory descriptor (8-row x 16-byte core matrices).  A
// stage covers 256 bytes of K (256 / SW samples); moving a stage on to its next samples means clearing one byte
// and setting one byte per (row, sample), so the fill work is ~24 byte stores per thread against 8
// tcgen05.mma.kind::i8 (128 x 256 x 32, u8 x u8 -> s32) per stage.  Accumulators live in tensor memory (256
// columns) and are read back once per tile.  Counts are exact integers (<= S).
//
// Warp roles (160 threads): warps 0-3 fill the stages and run the epilogue (warp w reads TMEM lanes 32w..32w+31),
// warp 4 allocates tensor memory and its lane 0 issues the MMAs.  full[st] (128 arrivals) hands a stage to the
// issuer; tcgen05.commit arrives on empty[st] when the MMAs that read the stage have completed.
#include "rc_common.cuh"

namespace {

constexpr int TC_M = 128, TC_N = 256, TC_KB = 256, TC_FILL = 128;
constexpr int TC_A_BYTES = TC_M * TC_KB, TC_B_BYTES = TC_N * TC_KB, TC_STAGE = TC_A_BYTES + TC_B_BYTES;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}
// UMMA shared-memory descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (rows 16 bytes apart);
// lbo = distance between the two 16-byte K chunks of one MMA, sbo = distance between 8-row groups.
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((saddr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor of kind::i8: D = s32 (bits 4-5 = 2), A and B unsigned 8 bit (0), both K-major, N >> 3 at
// bit 17, M >> 4 at bit 24
constexpr unsigned TC_IDESC = (2u << 4) | ((unsigned)(TC_N >> 3) << 17) | ((unsigned)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(a), "l"(b), "r"(TC_IDESC), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NS> __device__ __forceinline__ unsigned long long load_labels(const uint8_t* p) {
  if (NS == 8) return *reinterpret_cast<const unsigned long long*>(p);
  if (NS == 4) return *reinterpret_cast<const unsigned*>(p);
  return *reinterpret_cast<const unsigned short*>(p);
}

// one stage turn of one operand row: clear the bytes of the samples the stage held two turns ago, set the new ones
template <int SW, int ROWS>
__device__ __forceinline__ void flip_row(unsigned char* rowbase, unsigned long long oldw, unsigned long long neww) {
  constexpr int NS = TC_KB / SW;
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    const unsigned o = (unsigned)(oldw >> (8 * q)) & 0xffu, w = (unsigned)(neww >> (8 * q)) & 0xffu;
    if (o) { const unsigned kb = q * SW + o - 1; rowbase[(kb >> 4) * (ROWS * 16) + (kb & 15)] = 0; }
    if (w) { const unsigned kb = q * SW + w - 1; rowbase[(kb >> 4) * (ROWS * 16) + (kb & 15)] = 1; }
  }
}

// Lt: point-major labels [n][Rpad], 1-based, 0 = no sample (rows R..Rpad-1).  counts[i][j] for the tile and its mirror.
template <int SW>
__global__ void __launch_bounds__(160, 1) k_psm_tc(const uint8_t* __restrict__ Lt, long long n, long long Rpad, long long R,
                                                    int* __restrict__ counts) {
  constexpr int NS = TC_KB / SW;
  const long long i0 = (long long)blockIdx.y * TC_M, j0 = (long long)blockIdx.x * TC_N;
  if (j0 + TC_N <= i0) return;                              // tile entirely below the diagonal: covered by a mirror
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[2], empty[2], done;
  __shared__ unsigned tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int t = tid; t < 2 * TC_STAGE / 16; t += blockDim.x) reinterpret_cast<uint4*>(smem)[t] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mb_init(&full[0], TC_FILL); mb_init(&full[1], TC_FILL);
    mb_init(&empty[0], 1); mb_init(&empty[1], 1); mb_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_slot)), "r"(TC_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_slot;
  const long long nsteps = (R + NS - 1) / NS;

  if (warp < 4) {
    // ---- fill: thread t owns A row t and B rows t, 128 + t ----
    const bool okA = i0 + tid < n, okB0 = j0 + tid < n, okB1 = j0 + 128 + tid < n;
    const uint8_t* pA = Lt + (okA ? (i0 + tid) : 0) * Rpad;
    const uint8_t* pB0 = Lt + (okB0 ? (j0 + tid) : 0) * Rpad;
    const uint8_t* pB1 = Lt + (okB1 ? (j0 + 128 + tid) : 0) * Rpad;
    unsigned long long prev[2][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll 1
    for (long long t0 = 0; t0 < nsteps; t0 += 2) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const long long t = t0 + st;
        if (t < nsteps) {
          const unsigned long long wA = okA ? load_labels<NS>(pA + t * NS) : 0ull;
          const unsigned long long wB0 = okB0 ? load_labels<NS>(pB0 + t * NS) : 0ull;
          const unsigned long long wB1 = okB1 ? load_labels<NS>(pB1 + t * NS) : 0ull;
          if (t >= 2) mb_wait(&empty[st], (unsigned)(((t >> 1) - 1) & 1));
          unsigned char* sA = smem + st * TC_STAGE;
          unsigned char* sB = sA + TC_A_BYTES;
          flip_row<SW, TC_M>(sA + tid * 16, prev[st][0], wA);
          flip_row<SW, TC_N>(sB + tid * 16, prev[st][1], wB0);
          flip_row<SW, TC_N>(sB + (128 + tid) * 16, prev[st][2], wB1);
          prev[st][0] = wA; prev[st][1] = wB0; prev[st][2] = wB1;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mb_arrive(&full[st]);
        }
      }
    }
    // ---- epilogue: TMEM lane = tile row, column = tile column ----
    mb_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long i = i0 + warp * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < TC_N / 32; ++cb) {
      unsigned v[32];
      tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(cb * 32), v);
      const long long jb = j0 + cb * 32;
      if (jb >= n) break;                                   // warp-uniform
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const long long j = jb + q;
        if (i < n && j < n) {
          counts[i * n + j] = (int)v[q];
          counts[j * n + i] = (int)v[q];                    // mirror: consecutive lanes -> consecutive addresses
        }
      }
    }
  } else {
    // ---- MMA issue: one thread ----
    if (lane == 0) {
      const unsigned a0 = s32(smem), b0 = a0 + TC_A_BYTES;
      for (long long t = 0; t < nsteps; ++t) {
        const int st = (int)(t & 1);
        mb_wait(&full[st], (unsigned)((t >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int j = 0; j < TC_KB / 32; ++j) {
          const unsigned long long ad = umma_desc(a0 + st * TC_STAGE + j * (2 * TC_M * 16), TC_M * 16, 128);
          const unsigned long long bd = umma_desc(b0 + st * TC_STAGE + j * (2 * TC_N * 16), TC_N * 16, 128);
          umma_i8(tmem, ad, bd, (t > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(&empty[st]);
      }
      umma_commit(&done);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_N) : "memory");
  }
}

template <int SW>
int launch_tc(const uint8_t* Lt, int64_t n, int64_t Rpad, int64_t R, int* counts, cudaStream_t st) {
  const size_t smem = 2 * (size_t)TC_STAGE;
  RC_CUDA(cudaFuncSetAttribute(k_psm_tc<SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((n + TC_N - 1) / TC_N), (unsigned)((n + TC_M - 1) / TC_M));
  k_psm_tc<SW><<<grid, 160, smem, st>>>(Lt, (long long)n, (long long)Rpad, (long long)R, counts);
  RC_CUDA(cudaGetLastError());
  return RC_OK;
}

}  // namespace

// Tensor-core co-clustering counts; kmax = largest label present (<= 128).  Lt as produced by k_transpose in
// rc_post.cu (Rpad a multiple of 64, filler rows 0).  Asynchronous on `st`.
int rc_psm_counts_tc(const uint8_t* Lt, int64_t n, int64_t Rpad, int64_t R, int kmax, int* counts, cudaStream_t st) {
  if (kmax <= 32) return launch_tc<32>(Lt, n, Rpad, R, counts, st);
  if (kmax <= 64) return launch_tc<64>(Lt, n, Rpad, R, counts, st);
  if (kmax <= 128) return launch_tc<128>(Lt, n, Rpad, R, counts, st);
  rc_set_error("rc_psm_counts_tc: more than 128 labels");
  return RC_ERR_ARG;
}
// multiply-accumulates the tensor cores execute for one call (for the tensor-pipe utilisation figure)
double rc_psm_tc_macs(int64_t n, int64_t R, int kmax) {
  const int SW = kmax <= 32 ? 32 : kmax <= 64 ? 64 : 128;
  const int64_t NS = TC_KB / SW, nsteps = (R + NS - 1) / NS, tn = (n + TC_N - 1) / TC_N, tm = (n + TC_M - 1) / TC_M;
  int64_t tiles = 0;
  for (int64_t bi = 0; bi < tm; ++bi)
    for (int64_t bj = 0; bj < tn; ++bj) tiles += (bj * TC_N + TC_N > bi * TC_M);
  return (double)tiles * (double)nsteps * TC_M * TC_N * TC_KB;
}
