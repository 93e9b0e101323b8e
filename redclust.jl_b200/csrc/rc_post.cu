// rc_post.cu -- posterior similarity matrix and the minimum-posterior-expected-loss search.
//   PSM   sum(adjacencymatrix.(clusts)) ./ numsamples      /root/reference/src/mcmc.jl:560, src/utils.jl:59-63
//   MPEL  lossmatrix[i,j] = lossfn(c_i, c_j), argmin of column sums   /root/reference/src/pointestimate.jl:34-59
//         binder = randindex[3] (Mirkin), omARI = 1 - randindex[1], VI = varinfo, ID = max(H) - I
//         (Clustering.jl randindex / varinfo / mutualinfo, restated from their contingency-table definitions)
// The reference materialises S dense n x n Bool matrices; here the label matrix is transposed once
// (point-major, samples contiguous) and co-clustering counts are byte-compare popcounts -- exact integers.
#include <vector>
#include <algorithm>
#include "rc_common.cuh"

extern "C" const uint8_t* rc_sampler_dev_labels(const rc_sampler* s, int64_t* S, int64_t* n, int64_t* nchains, int* device);

namespace {

// Lt[i][r] = L[r][i]; rows r >= R are filled with 0 (every pair then "matches" Rpad - R extra times,
// which the count kernel subtracts).
__global__ void k_transpose(const uint8_t* __restrict__ L, int64_t R, int64_t n, int64_t Rpad, uint8_t* __restrict__ Lt) {
  __shared__ uint8_t t[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, i0 = (int64_t)blockIdx.x * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int64_t r = r0 + q, i = i0 + threadIdx.x;
    t[q][threadIdx.x] = (r < R && i < n) ? L[r * n + i] : 0;
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int64_t i = i0 + q, r = r0 + threadIdx.x;
    if (i < n && r < Rpad) Lt[i * Rpad + r] = t[threadIdx.x][q];
  }
}

__device__ __forceinline__ unsigned eq_bytes(unsigned a, unsigned b) {   // number of equal bytes (0..4)
  const unsigned x = a ^ b;
  const unsigned t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
  return __popc(~(t | x | 0x7f7f7f7fu));
}

// counts[i][j] = #{r < R : Lt[i][r] == Lt[j][r]}; 64 x 64 tile of pairs per CTA, upper triangle mirrored.
#define PT 64
#define PW 16   // 32-bit words (64 samples) per smem stage
__global__ void __launch_bounds__(256) k_psm_counts(const uint8_t* __restrict__ Lt, int64_t n, int64_t Rpad, int64_t R,
                                                    int* __restrict__ counts) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ unsigned A[PT][PW + 1], B[PT][PW + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  unsigned acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0;
  const int64_t i0 = (int64_t)bi * PT, j0 = (int64_t)bj * PT;
  const int64_t words = Rpad / 4;
  const unsigned* Lw = reinterpret_cast<const unsigned*>(Lt);
  for (int64_t w0 = 0; w0 < words; w0 += PW) {
    for (int t = threadIdx.x; t < PT * PW; t += 256) {
      const int row = t / PW, w = t % PW;
      const bool okw = w0 + w < words;
      A[row][w] = (okw && i0 + row < n) ? Lw[(i0 + row) * words + w0 + w] : 0u;
      B[row][w] = (okw && j0 + row < n) ? Lw[(j0 + row) * words + w0 + w] : 0x01010101u * 0xffu;   // never equal to A's filler
    }
    __syncthreads();
#pragma unroll 4
    for (int w = 0; w < PW; ++w) {
      unsigned a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = A[ty + 16 * q][w]; b[q] = B[tx + 16 * q][w]; }
#pragma unroll
      for (int qa = 0; qa < 4; ++qa)
#pragma unroll
        for (int qb = 0; qb < 4; ++qb) acc[qa][qb] += eq_bytes(a[qa], b[qb]);
    }
    __syncthreads();
  }
  const int pad = (int)(Rpad - R);
#pragma unroll
  for (int qa = 0; qa < 4; ++qa)
#pragma unroll
    for (int qb = 0; qb < 4; ++qb) {
      const int64_t i = i0 + ty + 16 * qa, j = j0 + tx + 16 * qb;
      if (i >= n || j >= n) continue;
      // words beyond `words` were filled with non-matching patterns; rows R..Rpad match for every pair
      const int v = (int)acc[qa][qb] - pad;
      counts[i * n + j] = v;
      counts[j * n + i] = v;
    }
}

__global__ void k_counts_to_psm(const int* __restrict__ counts, int64_t total, double denom, double* __restrict__ out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    out[t] = (double)counts[t] / denom;
}

int psm_counts_device(const uint8_t* L, int64_t R, int64_t n, int* counts) {
  const int64_t Rpad = (R + 63) & ~63LL;
  uint8_t* Lt = nullptr;
  RC_CUDA(cudaMalloc(&Lt, (size_t)n * Rpad));
  dim3 tb(32, 8), tg((unsigned)((n + 31) / 32), (unsigned)((Rpad + 31) / 32));
  k_transpose<<<tg, tb>>>(L, R, n, Rpad, Lt);
  const unsigned nb = (unsigned)((n + PT - 1) / PT);
  k_psm_counts<<<dim3(nb, nb), 256>>>(Lt, n, Rpad, R, counts);
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(Lt);
  RC_CUDA(e);
  return RC_OK;
}

// ---- MPEL ----------------------------------------------------------------------------------------
// One CTA per pair (i < j): contingency table in shared memory (packed 16-bit counters), then the loss.
__global__ void __launch_bounds__(256) k_pair_loss(const uint8_t* __restrict__ L, const int* __restrict__ Kc, int64_t S,
                                                   int64_t n, int loss, int wide, double* __restrict__ M) {
  const int64_t i = blockIdx.y, j = blockIdx.x;
  if (j <= i) return;
  extern __shared__ unsigned tab[];
  const int Ki = Kc[i], Kj = Kc[j];
  const int cells = Ki * Kj;
  const int nwords = wide ? cells : (cells + 1) / 2;
  for (int t = threadIdx.x; t < nwords; t += blockDim.x) tab[t] = 0;
  __syncthreads();
  const uint8_t* a = L + i * n; const uint8_t* b = L + j * n;
  for (int64_t x = threadIdx.x; x < n; x += blockDim.x) {
    const int cell = (int)(a[x] - 1) * Kj + (int)(b[x] - 1);
    if (wide) atomicAdd(&tab[cell], 1u);
    else atomicAdd(&tab[cell >> 1], (cell & 1) ? 0x10000u : 1u);
  }
  __syncthreads();
  // sums over the table: t2 = sum N^2, snl = sum N log N; margins via row / column passes
  double t2 = 0, snl = 0, nis = 0, njs = 0, hA = 0, hB = 0;
  const double dn = (double)n;
  for (int t = threadIdx.x; t < cells; t += blockDim.x) {
    const unsigned c = wide ? tab[t] : ((tab[t >> 1] >> ((t & 1) * 16)) & 0xffffu);
    if (c) { const double d = (double)c; t2 += d * d; snl += d * log(d); }
  }
  for (int r = threadIdx.x; r < Ki + Kj; r += blockDim.x) {
    unsigned s = 0;
    if (r < Ki) { for (int q = 0; q < Kj; ++q) { const int t = r * Kj + q; s += wide ? tab[t] : ((tab[t >> 1] >> ((t & 1) * 16)) & 0xffffu); } }
    else { const int q = r - Ki; for (int p = 0; p < Ki; ++p) { const int t = p * Kj + q; s += wide ? tab[t] : ((tab[t >> 1] >> ((t & 1) * 16)) & 0xffffu); } }
    if (s) {
      const double d = (double)s;
      if (r < Ki) { nis += d * d; hA += d * log(d); } else { njs += d * d; hB += d * log(d); }
    }
  }
  __shared__ double red[6][8];
  double v[6] = {t2, snl, nis, njs, hA, hB};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    for (int off = 16; off; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < 6; ++q) { double s = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[q][w]; v[q] = s; }
    t2 = v[0]; snl = v[1]; nis = v[2]; njs = v[3]; hA = v[4]; hB = v[5];
    double out;
    if (loss <= 1) {                                   // Clustering.randindex
      const double t1 = dn * (dn - 1) / 2, t3 = 0.5 * (nis + njs);
      const double nc = (dn * (dn * dn + 1) - (dn + 1) * nis - (dn + 1) * njs + 2 * (nis * njs) / dn) / (2 * (dn - 1));
      const double A = t1 + t2 - t3, Dg = -t2 + t3;
      if (loss == 0) out = Dg / t1;                    // Mirkin
      else out = 1 - ((t1 == nc) ? 0.0 : (A - nc) / (t1 - nc));
    } else {
      // H(A) = log n - (1/n) sum a log a;  I = (1/n) sum N log N - (1/n) sum a log a - (1/n) sum b log b + log n
      const double HA = log(dn) - hA / dn, HB = log(dn) - hB / dn;
      const double I = snl / dn - hA / dn - hB / dn + log(dn);
      out = loss == 2 ? (HA + HB - 2 * I) : ((HA > HB ? HA : HB) - I);
    }
    M[i * S + j] = out;
    M[j * S + i] = out;
  }
}

__global__ void k_colsum(const double* __restrict__ M, int64_t S, double* __restrict__ sums) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= S) return;
  double s = 0;
  for (int64_t i = 0; i < S; ++i) s += M[i * S + j];   // ascending-row order, as sum(lossmatrix, dims = 1)
  sums[j] = s;
}

}  // namespace

extern "C" {

int32_t rc_sampler_psm_counts_dev(const rc_sampler* s, int64_t chain0, int64_t nch, void* counts_dev) {
  if (!s || !counts_dev) { rc_set_error("rc_sampler_psm_counts_dev: null pointer"); return RC_ERR_ARG; }
  int64_t S, n, nchains; int device;
  const uint8_t* L = rc_sampler_dev_labels(s, &S, &n, &nchains, &device);
  if (chain0 < 0 || nch < 1 || chain0 + nch > nchains) { rc_set_error("rc_sampler_psm_counts_dev: bad chain range"); return RC_ERR_ARG; }
  RC_CUDA(cudaSetDevice(device));
  if (S == 0) { RC_CUDA(cudaMemset(counts_dev, 0, sizeof(int) * (size_t)n * n)); return RC_OK; }
  return psm_counts_device(L + (size_t)chain0 * S * n, nch * S, n, (int*)counts_dev);
}

int32_t rc_sampler_psm(const rc_sampler* s, int64_t chain0, int64_t nch, double* psm_out) {
  if (!s || !psm_out) { rc_set_error("rc_sampler_psm: null pointer"); return RC_ERR_ARG; }
  int64_t S, n, nchains; int device;
  rc_sampler_dev_labels(s, &S, &n, &nchains, &device);
  RC_CUDA(cudaSetDevice(device));
  int* counts = nullptr; double* out = nullptr;
  RC_CUDA(cudaMalloc(&counts, sizeof(int) * (size_t)n * n));
  int st = rc_sampler_psm_counts_dev(s, chain0, nch, counts);
  if (st) { cudaFree(counts); return st; }
  if (cudaMalloc(&out, sizeof(double) * (size_t)n * n) != cudaSuccess) { cudaFree(counts); rc_set_error("out of device memory"); return RC_ERR_CUDA; }
  k_counts_to_psm<<<148 * 8, 256>>>(counts, n * n, (double)(nch * S), out);   // ./ numsamples (0/0 = NaN if no samples)
  cudaError_t e = cudaMemcpy(psm_out, out, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost);
  cudaFree(counts); cudaFree(out);
  RC_CUDA(e);
  return RC_OK;
}

// first-appearance relabelling of host label vectors to 1..K (sortlabels, utils.jl:69-74) as bytes
static int compact_labels(const int64_t* labels, int64_t S, int64_t n, std::vector<uint8_t>& out, std::vector<int>& K) {
  out.resize((size_t)S * n); K.resize((size_t)S);
  std::vector<std::pair<int64_t, int64_t>> tmp((size_t)n);
  std::vector<int64_t> ids((size_t)n);
  for (int64_t s = 0; s < S; ++s) {
    const int64_t* l = labels + s * n;
    for (int64_t x = 0; x < n; ++x) tmp[x] = {l[x], x};
    std::sort(tmp.begin(), tmp.end());
    // first appearance position of every distinct label
    std::vector<std::pair<int64_t, int64_t>> firsts;   // (first position, label)
    for (int64_t x = 0; x < n; ++x) if (x == 0 || tmp[x].first != tmp[x - 1].first) firsts.push_back({tmp[x].second, tmp[x].first});
    std::sort(firsts.begin(), firsts.end());
    if (firsts.size() > 255) { rc_set_error("more than 255 clusters in sample %lld", (long long)s); return RC_ERR_SLOTS; }
    K[s] = (int)firsts.size();
    std::vector<std::pair<int64_t, int>> map;           // label -> id
    for (size_t q = 0; q < firsts.size(); ++q) map.push_back({firsts[q].second, (int)q + 1});
    std::sort(map.begin(), map.end());
    for (int64_t x = 0; x < n; ++x) {
      auto it = std::lower_bound(map.begin(), map.end(), std::make_pair(l[x], 0));
      out[s * n + x] = (uint8_t)it->second;
    }
  }
  return RC_OK;
}

int32_t rc_psm(const int64_t* labels, int64_t S, int64_t n, int32_t device, double* psm_out) {
  if (!labels || !psm_out || S < 1 || n < 1) { rc_set_error("rc_psm: null pointer or empty input"); return RC_ERR_ARG; }
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { rc_set_error("no CUDA device available (librcb200 has no CPU fallback)"); return RC_ERR_CUDA; }
  RC_CUDA(cudaSetDevice(device));
  std::vector<uint8_t> L; std::vector<int> K;
  int st = compact_labels(labels, S, n, L, K);
  if (st) return st;
  uint8_t* dL = nullptr; int* counts = nullptr; double* out = nullptr;
  RC_CUDA(cudaMalloc(&dL, L.size()));
  RC_CUDA(cudaMemcpy(dL, L.data(), L.size(), cudaMemcpyHostToDevice));
  RC_CUDA(cudaMalloc(&counts, sizeof(int) * (size_t)n * n));
  st = psm_counts_device(dL, S, n, counts);
  if (!st) {
    RC_CUDA(cudaMalloc(&out, sizeof(double) * (size_t)n * n));
    k_counts_to_psm<<<148 * 8, 256>>>(counts, n * n, (double)S, out);
    cudaError_t e = cudaMemcpy(psm_out, out, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { rc_set_error("rc_psm: %s", cudaGetErrorString(e)); st = RC_ERR_CUDA; }
  }
  cudaFree(dL); cudaFree(counts); cudaFree(out);
  return st;
}

int32_t rc_mpel(const int64_t* labels, int64_t S, int64_t n, int32_t loss, int32_t device, double* loss_sums, int64_t* best) {
  if (!labels || S < 1 || n < 1 || loss < 0 || loss > 3) { rc_set_error("rc_mpel: bad arguments"); return RC_ERR_ARG; }
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { rc_set_error("no CUDA device available (librcb200 has no CPU fallback)"); return RC_ERR_CUDA; }
  RC_CUDA(cudaSetDevice(device));
  std::vector<uint8_t> L; std::vector<int> K;
  int st = compact_labels(labels, S, n, L, K);
  if (st) return st;
  int kmax = 0;
  for (int k : K) kmax = std::max(kmax, k);
  const int wide = n > 65535;
  const size_t smem = (size_t)kmax * kmax * (wide ? 4 : 2) + 8;
  if (smem > 200 * 1024) { rc_set_error("rc_mpel: contingency table of %d x %d clusters does not fit shared memory", kmax, kmax); return RC_ERR_SLOTS; }
  uint8_t* dL = nullptr; int* dK = nullptr; double *M = nullptr, *sums = nullptr;
  RC_CUDA(cudaMalloc(&dL, L.size()));
  RC_CUDA(cudaMalloc(&dK, sizeof(int) * S));
  RC_CUDA(cudaMalloc(&M, sizeof(double) * (size_t)S * S));
  RC_CUDA(cudaMalloc(&sums, sizeof(double) * S));
  RC_CUDA(cudaMemcpy(dL, L.data(), L.size(), cudaMemcpyHostToDevice));
  RC_CUDA(cudaMemcpy(dK, K.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
  RC_CUDA(cudaMemset(M, 0, sizeof(double) * (size_t)S * S));
  cudaFuncSetAttribute(k_pair_loss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_pair_loss<<<dim3((unsigned)S, (unsigned)S), 256, smem>>>(dL, dK, S, n, loss, wide, M);
  k_colsum<<<(unsigned)((S + 127) / 128), 128>>>(M, S, sums);
  std::vector<double> hs((size_t)S);
  cudaError_t e = cudaMemcpy(hs.data(), sums, sizeof(double) * S, cudaMemcpyDeviceToHost);
  cudaFree(dL); cudaFree(dK); cudaFree(M); cudaFree(sums);
  RC_CUDA(e);
  int64_t b = 0;
  for (int64_t i = 1; i < S; ++i) if (hs[i] < hs[b]) b = i;     // argmin: first minimum
  if (loss_sums) memcpy(loss_sums, hs.data(), sizeof(double) * S);
  if (best) *best = b;
  return RC_OK;
}

}  // extern "C"
