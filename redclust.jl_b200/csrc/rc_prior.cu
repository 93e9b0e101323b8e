// rc_prior.cu -- the O(n^2) pieces of fitprior on the device-resident dissimilarity matrix.
//   k-medoids      Clustering.kmedoids(dissM, k; maxiter) as called by fitprior  /root/reference/src/prior.jl:55-71
//                  and by runsampler's default initialisation                     /root/reference/src/mcmc.jl:519-527
//   pair statistics  A = uppertriangle(dissM)[adjacency], B = the rest; fit_mle(Gamma, .) needs count, sum and
//                  sum of logs of each                                            /root/reference/src/prior.jl:73-110
// Sums run over the exact fixed-point images (Dq, Lq), so every comparison and every statistic is independent of the
// order of summation: a host restatement with the same integers reproduces the medoids bit for bit.
#include <vector>
#include <cstring>
#include "rc_common.cuh"

namespace {

// assignment step: nearest medoid (first minimum, as argmin over the medoid rows)
__global__ void k_km_assign(const double* __restrict__ D, int64_t n, const int* __restrict__ med, int k, int* __restrict__ assign,
                            int* __restrict__ changed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double best = D[(size_t)med[0] * n + i];
  int arg = 0;
  for (int c = 1; c < k; ++c) {
    const double v = D[(size_t)med[c] * n + i];      // row of the medoid: consecutive threads, consecutive addresses
    if (v < best) { best = v; arg = c; }
  }
  if (assign[i] != arg) { assign[i] = arg; *changed = 1; }
}

// s[i] = sum of Dq[i][j] over the members j of i's cluster (one CTA per row, grid-stride over rows)
__global__ void __launch_bounds__(256) k_km_rowsum(const longlong2* __restrict__ DL, int64_t n, const int* __restrict__ assign,
                                                   long long* __restrict__ s) {
  __shared__ long long red[8];
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const longlong2* row = DL + (size_t)i * n;
    const int li = assign[i];
    long long acc = 0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x)
      if (assign[j] == li) acc += row[j].x;
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int w = 0; w < 8; ++w) t += red[w];
      s[i] = t;
    }
    __syncthreads();
  }
}
// medoid update: the member with the smallest sum, lowest index on ties (two passes of integer atomics)
__global__ void k_km_min1(const long long* __restrict__ s, const int* __restrict__ assign, int64_t n, long long* __restrict__ mins) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) atomicMin(&mins[assign[i]], s[i]);
}
__global__ void k_km_min2(const long long* __restrict__ s, const int* __restrict__ assign, int64_t n, const long long* __restrict__ mins,
                          int* __restrict__ newmed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && s[i] == mins[assign[i]]) atomicMin(&newmed[assign[i]], (int)i);
}
__global__ void k_km_finish(const int* __restrict__ med, int k, int* __restrict__ newmed, int* __restrict__ changed) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  if (newmed[c] == 0x7f7f7f7f) newmed[c] = med[c];          // empty cluster keeps its medoid
  if (newmed[c] != med[c]) *changed = 1;
}
// total cost as an exact integer: per-block partial sums of Dq[medoid(i)][i]
__global__ void __launch_bounds__(256) k_km_cost(const longlong2* __restrict__ DL, int64_t n, const int* __restrict__ med,
                                                 const int* __restrict__ assign, long long* __restrict__ partial) {
  long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += DL[(size_t)med[assign[i]] * n + i].x;
  __shared__ long long red[8];
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { long long t = 0; for (int w = 0; w < 8; ++w) t += red[w]; partial[blockIdx.x] = t; }
}

// row i: sums over j > i of (Dq, Lq) within i's cluster and over all j > i, and the within count
__global__ void __launch_bounds__(256) k_pair_rows(const longlong2* __restrict__ DL, int64_t n, const int* __restrict__ lab,
                                                   long long* __restrict__ out) {
  __shared__ long long red[5][8];
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const longlong2* row = DL + (size_t)i * n;
    const int li = lab[i];
    long long v[5] = {0, 0, 0, 0, 0};          // within D, within L, all D, all L, within count
    for (int64_t j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const longlong2 e = row[j];
      v[2] += e.x; v[3] += e.y;
      if (lab[j] == li) { v[0] += e.x; v[1] += e.y; v[4] += 1; }
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      for (int off = 16; off; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
      if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      long long t = 0;
      for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
      out[(size_t)i * 5 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

struct Scratch {      // frees everything on scope exit
  std::vector<void*> p;
  ~Scratch() { for (void* q : p) cudaFree(q); }
  template <class T> cudaError_t get(T** out, size_t count) {
    cudaError_t e = cudaMalloc((void**)out, sizeof(T) * (count ? count : 1));
    if (e == cudaSuccess) p.push_back(*out);
    return e;
  }
};

}  // namespace

extern "C" {

int32_t rc_data_copy_row(const rc_data* d, int64_t i, double* row_out) {
  if (!d || !row_out || i < 0 || i >= d->n) { rc_set_error("rc_data_copy_row: bad arguments"); return RC_ERR_ARG; }
  RC_CUDA(cudaSetDevice(d->device));
  RC_CUDA(cudaMemcpy(row_out, d->D + (size_t)i * d->n, sizeof(double) * d->n, cudaMemcpyDeviceToHost));
  return RC_OK;
}

int32_t rc_kmedoids(const rc_data* d, int64_t k, const int64_t* init_medoids, int64_t maxiter, int64_t* assignments,
                    int64_t* medoids, double* totalcost, int32_t* converged, int64_t* iterations) {
  if (!d || !init_medoids || !assignments || !medoids || k < 1 || k > d->n || maxiter < 0) {
    rc_set_error("rc_kmedoids: bad arguments (need 1 <= k <= n)"); return RC_ERR_ARG;
  }
  const int64_t n = d->n;
  RC_CUDA(cudaSetDevice(d->device));
  std::vector<int> hmed((size_t)k);
  for (int64_t c = 0; c < k; ++c) {
    if (init_medoids[c] < 0 || init_medoids[c] >= n) { rc_set_error("rc_kmedoids: initial medoid out of range"); return RC_ERR_ARG; }
    hmed[c] = (int)init_medoids[c];
  }
  Scratch S;
  int *med, *newmed, *assign, *flag; long long *s, *mins, *partial;
  const int nb = 148 * 4;
  if (S.get(&med, k) || S.get(&newmed, k) || S.get(&assign, n) || S.get(&flag, 2) || S.get(&s, n) || S.get(&mins, k) || S.get(&partial, nb)) {
    rc_set_error("rc_kmedoids: out of device memory"); return RC_ERR_CUDA;
  }
  RC_CUDA(cudaMemcpy(med, hmed.data(), sizeof(int) * k, cudaMemcpyHostToDevice));
  RC_CUDA(cudaMemset(assign, 0xff, sizeof(int) * n));
  const unsigned gn = (unsigned)((n + 255) / 256), gk = (unsigned)((k + 255) / 256);
  k_km_assign<<<gn, 256>>>(d->D, n, med, (int)k, assign, flag);
  bool conv = false;
  int64_t it = 0;
  for (; it < maxiter; ++it) {
    RC_CUDA(cudaMemsetAsync(flag, 0, 2 * sizeof(int)));
    RC_CUDA(cudaMemsetAsync(mins, 0x7f, sizeof(long long) * k));          // 0x7f7f... > any row sum (< 2^61)
    RC_CUDA(cudaMemsetAsync(newmed, 0x7f, sizeof(int) * k));              // 0x7f7f7f7f > any index: start value of the atomicMin
    k_km_rowsum<<<148 * 8, 256>>>(d->DL, n, assign, s);
    k_km_min1<<<gn, 256>>>(s, assign, n, mins);
    k_km_min2<<<gn, 256>>>(s, assign, n, mins, newmed);
    k_km_finish<<<gk, 256>>>(med, (int)k, newmed, flag);
    k_km_assign<<<gn, 256>>>(d->D, n, newmed, (int)k, assign, flag + 1);
    int h[2] = {0, 0};
    RC_CUDA(cudaMemcpy(h, flag, sizeof(h), cudaMemcpyDeviceToHost));
    std::swap(med, newmed);
    if (!h[0] && !h[1]) { conv = true; break; }
  }
  k_km_cost<<<nb, 256>>>(d->DL, n, med, assign, partial);
  std::vector<long long> hp((size_t)nb);
  std::vector<int> ha((size_t)n);
  RC_CUDA(cudaMemcpy(hp.data(), partial, sizeof(long long) * nb, cudaMemcpyDeviceToHost));
  RC_CUDA(cudaMemcpy(ha.data(), assign, sizeof(int) * n, cudaMemcpyDeviceToHost));
  RC_CUDA(cudaMemcpy(hmed.data(), med, sizeof(int) * k, cudaMemcpyDeviceToHost));
  rc_i128 tot = {0ull, 0ll};
  for (long long v : hp) rc_add128(tot, v);
  for (int64_t i = 0; i < n; ++i) assignments[i] = (int64_t)ha[i] + 1;
  for (int64_t c = 0; c < k; ++c) medoids[c] = hmed[c];
  if (totalcost) *totalcost = ((double)tot.hi * 18446744073709551616.0 + (double)tot.lo) / (double)(1ull << d->qD);
  if (converged) *converged = conv ? 1 : 0;
  if (iterations) *iterations = it + (conv ? 1 : 0);
  return RC_OK;
}

int32_t rc_pair_stats(const rc_data* d, const int64_t* labels, int64_t* rows_out) {
  if (!d || !labels || !rows_out) { rc_set_error("rc_pair_stats: null pointer"); return RC_ERR_ARG; }
  const int64_t n = d->n;
  RC_CUDA(cudaSetDevice(d->device));
  std::vector<int> hl((size_t)n);
  for (int64_t i = 0; i < n; ++i) hl[i] = (int)labels[i];
  Scratch S;
  int* lab; long long* out;
  if (S.get(&lab, n) || S.get(&out, 5 * n)) { rc_set_error("rc_pair_stats: out of device memory"); return RC_ERR_CUDA; }
  RC_CUDA(cudaMemcpy(lab, hl.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  k_pair_rows<<<148 * 8, 256>>>(d->DL, n, lab, out);
  RC_CUDA(cudaMemcpy(rows_out, out, sizeof(long long) * 5 * n, cudaMemcpyDeviceToHost));
  return RC_OK;
}

}  // extern "C"
