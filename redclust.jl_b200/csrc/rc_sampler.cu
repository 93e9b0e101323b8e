// rc_sampler.cu -- the persistent chain kernel: one launch runs every iteration of runsampler's loop
// (/root/reference/src/mcmc.jl:537-555) for every chain on the device.
//
//   sample_r!  (mcmc.jl:80-136)      -> update_r()
//   sample_p!  (mcmc.jl:138-155)     -> update_p()
//   sample_labels! (mcmc.jl:356-479) -> splitmerge_step() x numMH, then full_scan()
//   sample_labels_Gibbs! (:158-256)  -> full_scan(): reduce_row() + scan_decide()
//   sample_labels_Gibbs_restricted! (:259-354) -> restricted_scan()
//   loglik (:1-56), logprior (:58-78), record step (:546-554), sortlabels (utils.jl:69-74)
//
// Data layout / algorithm (DESIGN.md sections 2-4):
//   * DL[i][j] = {Dq, Lq}: 64-bit fixed-point images of D and log D.  Every cluster sum is an exact
//     integer, so results do not depend on tiling, lane count or reduction order: the kernel is
//     bit-identical to the CPU oracle by construction.
//   * reduce_row(x): s_k = sum_{j in k} DL[x][j] for all slots k at once.  Columns are kept in a
//     (tile, label)-sorted permutation whose label runs are padded to groups of 8; a lane sums one
//     group, a warp-shuffle segmented scan combines the groups of a run, run tails accumulate into
//     per-warp bins.  No atomics, no label compares in the inner loop.
//   * The K x K block-sum matrices W (128-bit integers) are maintained incrementally from the row
//     sums of every accepted move, so loglik() never re-reads D: it is O(K^2) transcendentals.
//   * Split-merge: row sums of the members of ci u cj are taken once (launch state); the restricted
//     scans then only gather the |ci u cj| entries of one row per step.
#include "rc_sampler.cuh"

namespace {

struct Scal {
  double r, p, logp, log1mp;
  double ltp;
  double ll_cur, ll_fin;
  double dtmp[4];
  rc_i128 aaD, aaL, abD, abL, bbD, bbL;   // block sums of the proposed state (rows a, b)
  int K;
  int status;
  int rebuild;
  int fslotA, fslotB;                     // slots whose W rows come from the scratch rows (-1: none)
  int itmp[8];
};

struct Ctx {
  int n, cap, tiles;
  int qD, qL;
  const longlong2* DL;
  const rc_kparams* kp;
  // shared memory
  uint8_t* lab;
  uint8_t* labL;
  unsigned short* perm;
  uint8_t* glabel;
  unsigned short* runStart;
  unsigned int* cnt;
  int* tileStart;
  longlong2* partial;     // [RC_NWARP][cap]; aliased by rowA/rowB during loglik of a proposed state
  int* sizes;
  int* szL;
  int* itmp;              // [cap]
  uint8_t* clist;         // [cap]
  long long* red;         // [RC_NWARP * 4]
  Scal* sc;
  // per-chain global memory
  rc_i128* WD;
  rc_i128* WL;
  longlong2* T;
  unsigned short* Slist;
  double* terms;
  unsigned long long key;
};

__device__ __forceinline__ int tri(int k, int t, int cap) { return k < t ? k * cap + t : t * cap + k; }

__device__ __forceinline__ long long shfl_up_ll(long long v, int off) { return __shfl_up_sync(0xffffffffu, v, off); }
__device__ __forceinline__ long long shfl_xor_ll(long long v, int off) { return __shfl_xor_sync(0xffffffffu, v, off); }

// ------------------------------------------------------------------------------------------------
// (tile, label)-sorted column permutation with label runs padded to multiples of RC_GROUP.
// ------------------------------------------------------------------------------------------------
__device__ void build_perm(const Ctx& c, const uint8_t* lab) {
  const int tid = threadIdx.x, E = c.tiles * c.cap;
  for (int t = tid; t < E; t += RC_NTHR) c.cnt[t] = 0;
  __syncthreads();
  for (int j = tid; j < c.n; j += RC_NTHR) atomicAdd(&c.cnt[(j >> RC_LOGW) * c.cap + lab[j]], 1u);
  __syncthreads();
  if (tid < 32) {
    const int chunk = (E + 31) / 32;
    const int b = tid * chunk, e = min(E, b + chunk);
    unsigned s = 0;
    for (int t = b; t < e; ++t) s += (c.cnt[t] + 7u) & ~7u;
    unsigned incl = s;
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
      if (tid >= off) incl += o;
    }
    unsigned run = incl - s;
    for (int t = b; t < e; ++t) { c.runStart[t] = (unsigned short)run; run += (c.cnt[t] + 7u) & ~7u; }
    if (tid == 31) c.runStart[E] = (unsigned short)incl;
  }
  __syncthreads();
  const int total = c.runStart[E];
  for (int t = tid; t <= c.tiles; t += RC_NTHR) c.tileStart[t] = c.runStart[t == c.tiles ? E : t * c.cap] >> 3;
  for (int q = tid; q < total; q += RC_NTHR) c.perm[q] = (unsigned short)RC_DUMMY;
  __syncthreads();   // cnt is still being read above by nobody, but keep phases separate for clarity
  for (int t = tid; t < E; t += RC_NTHR) {
    const int g0 = c.runStart[t] >> 3, g1 = c.runStart[t + 1] >> 3;
    const uint8_t l = (uint8_t)(t % c.cap);
    for (int g = g0; g < g1; ++g) c.glabel[g] = l;
    c.cnt[t] = 0;
  }
  __syncthreads();
  for (int j = tid; j < c.n; j += RC_NTHR) {
    const int e = (j >> RC_LOGW) * c.cap + lab[j];
    const unsigned pos = c.runStart[e] + atomicAdd(&c.cnt[e], 1u);
    c.perm[pos] = (unsigned short)(j & (RC_W - 1));
  }
  __syncthreads();
}

// Point j (column) moved from slot a to slot b: patch the permutation in place (warp 0).  If the run of
// (tile, b) has no free padding entry the caller rebuilds.
__device__ void patch_perm(const Ctx& c, int j, int a, int b) {
  const int lane = threadIdx.x & 31;
  const int tile = j >> RC_LOGW;
  const unsigned short idx = (unsigned short)(j & (RC_W - 1));
  {
    const int e = tile * c.cap + a;
    const int p0 = c.runStart[e], p1 = c.runStart[e + 1];
    for (int pb = p0; pb < p1; pb += 32) {
      const int p = pb + lane;
      if (p < p1 && c.perm[p] == idx) c.perm[p] = (unsigned short)RC_DUMMY;
    }
  }
  __syncwarp();
  bool done = false;
  {
    const int e = tile * c.cap + b;
    const int p0 = c.runStart[e], p1 = c.runStart[e + 1];
    for (int pb = p0; pb < p1 && !done; pb += 32) {
      const int p = pb + lane;
      const unsigned m = __ballot_sync(0xffffffffu, p < p1 && c.perm[p] == (unsigned short)RC_DUMMY);
      if (m) {
        if (lane == __ffs(m) - 1) c.perm[p] = idx;
        done = true;
      }
    }
  }
  if (!done && lane == 0) c.sc->rebuild = 1;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// reduce_row: partial[warp][slot] = sums of DL[x][j] over the columns j of each label run handled by
// the warp.  The caller __syncthreads() and adds the RC_NWARP partials.   (matsum(D,[i],clust_k) and
// matsum(logD,[i],clust_k) for every k at once: mcmc.jl:210-213, 311-318; utils.jl:9-17.)
// ------------------------------------------------------------------------------------------------
__device__ void reduce_row(const Ctx& c, int x) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  longlong2* part = c.partial + warp * c.cap;
  for (int s = lane; s < c.cap; s += 32) part[s] = make_longlong2(0, 0);
  __syncwarp();
  const longlong2* row = c.DL + (size_t)x * c.n;
  for (int tile = 0; tile < c.tiles; ++tile) {
    const int g0 = c.tileStart[tile], g1 = c.tileStart[tile + 1];
    const longlong2* rt = row + tile * RC_W;
    for (int gb = g0 + warp * 32; gb < g1; gb += RC_NWARP * 32) {
      const int g = gb + lane;
      const bool valid = g < g1;
      const int lab = valid ? (int)c.glabel[g] : 0x100;
      long long d = 0, l = 0;
      if (valid) {
        const uint4 pk = *reinterpret_cast<const uint4*>(c.perm + g * RC_GROUP);
        const unsigned w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int e = 0; e < RC_GROUP; ++e) {
          const unsigned idx = (w[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
          if (idx != RC_DUMMY) {
            const longlong2 v = __ldg(rt + idx);
            d += v.x; l += v.y;
          }
        }
      }
      const int prev = __shfl_up_sync(0xffffffffu, lab, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != lab);
      const int runstart = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
      const int pos = lane - runstart;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const long long od = shfl_up_ll(d, off), ol = shfl_up_ll(l, off);
        if (pos >= off) { d += od; l += ol; }
      }
      const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
      if (valid && tail) {
        longlong2 a = part[lab];
        a.x += d; a.y += l;
        part[lab] = a;
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ longlong2 bin_total(const Ctx& c, int s) {
  longlong2 a = c.partial[s];
#pragma unroll
  for (int w = 1; w < RC_NWARP; ++w) {
    const longlong2 b = c.partial[w * c.cap + s];
    a.x += b.x; a.y += b.y;
  }
  return a;
}

// ------------------------------------------------------------------------------------------------
// One step of the full Gibbs scan for point i (mcmc.jl:192-253), executed by warp 0 after reduce_row(i).
// Lane l owns slots l, l+32, l+64, l+96.
// ------------------------------------------------------------------------------------------------
__device__ void scan_decide(const Ctx& c, int i, unsigned it) {
  const int lane = threadIdx.x & 31;
  const rc_kparams& kp = *c.kp;
  const rc_params& P = kp.P;
  const int cap = c.cap;
  const int li = c.lab[i];
  const longlong2 self = __ldg(c.DL + (size_t)i * c.n + i);
  const double r = c.sc->r, logp = c.sc->logp, log1mp = c.sc->log1mp;

  long long bd[RC_NS], bl[RC_NS];
  int sz[RC_NS];
  unsigned occ[RC_NS];
#pragma unroll
  for (int w = 0; w < RC_NS; ++w) {
    const int s = w * 32 + lane;
    bd[w] = 0; bl[w] = 0; sz[w] = 0;
    if (s < cap) {
      const longlong2 t = bin_total(c, s);
      bd[w] = t.x; bl[w] = t.y; sz[w] = c.sizes[s];
      if (s == li) { bd[w] -= self.x; bl[w] -= self.y; sz[w] -= 1; }       // :193-194 detach i
    }
    occ[w] = __ballot_sync(0xffffffffu, sz[w] > 0);
  }
  int Ki = 0, e = -1;
#pragma unroll
  for (int w = 0; w < RC_NS; ++w) {
    Ki += __popc(occ[w]);
    const int lim = cap - w * 32;
    const unsigned capmask = lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
    const unsigned emp = ~occ[w] & capmask;
    if (e < 0 && emp) e = w * 32 + __ffs(emp) - 1;                          // findfirst(clustsizes .== 0)
  }
  const bool hasnew = (P.maxK == 0 || Ki < P.maxK) && Ki < c.n;             // :198
  if (hasnew && e < 0) {                                                    // slot capacity exhausted
    if (lane == 0) c.sc->status = RC_ERR_SLOTS;
    return;
  }
  // per-slot terms (:206-242)
  double L1[RC_NS], lpr[RC_NS], L2p[RC_NS];
  double acc = 0.0;
#pragma unroll
  for (int w = 0; w < RC_NS; ++w) {
    L1[w] = 0.0; lpr[w] = 0.0; L2p[w] = 0.0;
    if (sz[w] > 0) {
      const double szd = (double)sz[w];
      const double sD = rc_dequant(bd[w], c.qD), sL = rc_dequant(bl[w], c.qL);
      const double a_i = P.alpha + P.delta1 * szd, b_i = P.beta + sD;
      const double z_i = P.zeta + P.delta2 * szd, g_i = P.gamma + sD;
      L1[w] = kp.LGA[sz[w]] + kp.abratio - a_i * rc_log(b_i) + (P.delta1 - 1) * sL - szd * kp.lgd1;
      lpr[w] = kp.LOGN[sz[w] + 1] + logp + rc_log((double)(sz[w] - 1) + r) - kp.LOGN[sz[w]];
      L2p[w] = kp.LGZ[sz[w]] - z_i * rc_log(g_i) + kp.zgratio + (P.delta2 - 1) * sL - szd * kp.lgd2;
      acc += L2p[w];                                                        // vecsum: lane-wise ascending slots
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);   // :243 canonical butterfly
  const double L2i = acc;
  // log-probabilities, Gumbel-max (utils.jl:2-6)
  double lp[RC_NS];
  int kk[RC_NS];
  bool have[RC_NS];
  int base = 0;
  bool anynan = false;
  double mn = RC_INF;
#pragma unroll
  for (int w = 0; w < RC_NS; ++w) {
    have[w] = false; lp[w] = 0.0; kk[w] = 0;
    const int s = w * 32 + lane;
    if (sz[w] > 0) {
      const double L2 = L2i - L2p[w];
      lp[w] = lpr[w] + (L1[w] + (P.repulsion ? L2 : copysign(0.0, L2)));
      kk[w] = base + __popc(occ[w] & ((1u << lane) - 1u));
      have[w] = true;
    } else if (hasnew && s == e) {                                          // :228-230 new cluster
      const double L2 = L2i - 0.0;
      lp[w] = (kp.LOGN[Ki + 1] + r * log1mp) + (0.0 + (P.repulsion ? L2 : copysign(0.0, L2)));
      kk[w] = Ki;
      have[w] = true;
    }
    if (have[w]) {
      if (rc_isnan(lp[w])) anynan = true;
      else if (lp[w] < mn) mn = lp[w];
    }
    base += __popc(occ[w]);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, mn, off);
    if (o < mn) mn = o;
  }
  anynan = __any_sync(0xffffffffu, anynan);
  if (anynan) mn = RC_NAN;                                                  // Julia minimum propagates NaN
  // argmax of gumbel + shifted logprob; NaN is maximal, first index wins ties
  double bg = 0.0; int bk = 0x7fffffff, bs = -1; bool bnan = false;
#pragma unroll
  for (int w = 0; w < RC_NS; ++w) {
    if (!have[w]) continue;
    const rc_draw dr = rc_draw2(c.key, it, RC_SITE_SCAN, 0, (uint32_t)i, (uint32_t)(kk[w] >> 1));
    const double u = (kk[w] & 1) ? dr.u1 : dr.u0;
    const double g = -rc_log(-rc_log(u)) + (lp[w] - mn);
    const bool gn = rc_isnan(g);
    bool better;
    if (bs < 0) better = true;
    else if (gn) better = !bnan || kk[w] < bk;
    else if (bnan) better = false;
    else better = g > bg || (g == bg && kk[w] < bk);
    if (better) { bg = g; bk = kk[w]; bs = w * 32 + lane; bnan = gn; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double og = __shfl_xor_sync(0xffffffffu, bg, off);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
    const int os = __shfl_xor_sync(0xffffffffu, bs, off);
    const int on = __shfl_xor_sync(0xffffffffu, (int)bnan, off);
    bool better;
    if (os < 0) better = false;
    else if (bs < 0) better = true;
    else if (on) better = !bnan || ok < bk;
    else if (bnan) better = false;
    else better = og > bg || (og == bg && ok < bk);
    if (better) { bg = og; bk = ok; bs = os; bnan = on != 0; }
  }
  const int cnew = bs;
  if (lane == 0) {                                                          // :250-252
    c.lab[i] = (uint8_t)cnew;
    c.sizes[li] -= 1;
    c.sizes[cnew] += 1;
  }
  if (cnew == li) return;
  // the point moved: update the block-sum matrices from its row sums (exact integers)
  {
    const int a = li, b = cnew;
#pragma unroll
    for (int w = 0; w < RC_NS; ++w) {
      const int s = w * 32 + lane;
      if (s >= cap) continue;
      if (s == a) {
        const int ix = tri(a, a, cap);
        rc_i128 x = c.WD[ix]; rc_sub128(x, rc_make128(2 * bd[w] + self.x)); c.WD[ix] = x;
        rc_i128 y = c.WL[ix]; rc_sub128(y, rc_make128(2 * bl[w] + self.y)); c.WL[ix] = y;
      } else if (bd[w] != 0 || bl[w] != 0) {
        const int ix = tri(a, s, cap);
        rc_i128 x = c.WD[ix]; rc_sub128(x, rc_make128(bd[w])); c.WD[ix] = x;
        rc_i128 y = c.WL[ix]; rc_sub128(y, rc_make128(bl[w])); c.WL[ix] = y;
      }
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < RC_NS; ++w) {
      const int s = w * 32 + lane;
      if (s >= cap) continue;
      if (s == b) {
        const int ix = tri(b, b, cap);
        rc_i128 x = c.WD[ix]; rc_add128(x, rc_make128(2 * bd[w] + self.x)); c.WD[ix] = x;
        rc_i128 y = c.WL[ix]; rc_add128(y, rc_make128(2 * bl[w] + self.y)); c.WL[ix] = y;
      } else if (bd[w] != 0 || bl[w] != 0) {
        const int ix = tri(b, s, cap);
        rc_i128 x = c.WD[ix]; rc_add128(x, rc_make128(bd[w])); c.WD[ix] = x;
        rc_i128 y = c.WL[ix]; rc_add128(y, rc_make128(bl[w])); c.WL[ix] = y;
      }
    }
    __syncwarp();
    patch_perm(c, i, a, b);
  }
}

// sample_labels_Gibbs! (mcmc.jl:158-256) on the chain's state.
__device__ void full_scan(const Ctx& c, unsigned it) {
  for (int i = 0; i < c.n; ++i) {
    reduce_row(c, i);
    __syncthreads();
    if (threadIdx.x < 32) scan_decide(c, i, it);
    __syncthreads();
    if (c.sc->status) return;
    if (c.sc->rebuild) {
      __syncthreads();
      if (threadIdx.x == 0) c.sc->rebuild = 0;
      build_perm(c, c.lab);
    }
  }
  if (threadIdx.x < 32) {                                                    // :254
    int K = 0;
    for (int s = threadIdx.x; s < c.cap; s += 32) K += c.sizes[s] > 0;
    for (int off = 16; off; off >>= 1) K += __shfl_xor_sync(0xffffffffu, K, off);
    if (threadIdx.x == 0) c.sc->K = K;
  }
  __syncthreads();
}

// Block sums from scratch: W[k][t] = sum_{x in k, y in t} DL[x][y]  (n row reductions).
__device__ void init_W(const Ctx& c) {
  for (int t = threadIdx.x; t < c.cap * c.cap; t += RC_NTHR) {
    rc_i128 z; z.lo = 0; z.hi = 0;
    c.WD[t] = z; c.WL[t] = z;
  }
  __syncthreads();
  for (int x = 0; x < c.n; ++x) {
    reduce_row(c, x);
    __syncthreads();
    const int k = c.lab[x];
    for (int t = k + (int)threadIdx.x; t < c.cap; t += RC_NTHR) {
      const longlong2 b = bin_total(c, t);
      if (b.x != 0 || b.y != 0) {
        const int ix = k * c.cap + t;
        rc_i128 a = c.WD[ix]; rc_add128(a, b.x); c.WD[ix] = a;
        rc_i128 d = c.WL[ix]; rc_add128(d, b.y); c.WL[ix] = d;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// loglik (mcmc.jl:1-56) from the block sums.  `sz` are the cluster sizes of the evaluated state; rows of
// slots fslotA / fslotB (proposed state of a split-merge step) come from the scratch rows aliased on
// c.partial, everything else from the chain's W.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ rc_i128 getW(const Ctx& c, bool logm, int k, int t) {
  const Scal& s = *c.sc;
  const int A = s.fslotA, B = s.fslotB;
  if (k == A || t == A || k == B || t == B) {
    if (A >= 0 && B >= 0 && ((k == A && t == B) || (k == B && t == A))) return logm ? s.abL : s.abD;
    if (k == A && t == A) return logm ? s.aaL : s.aaD;
    if (k == B && t == B) return logm ? s.bbL : s.bbD;
    const rc_i128* rows = reinterpret_cast<const rc_i128*>(c.partial);    // [rowA_D | rowA_L | rowB_D | rowB_L] x cap
    const bool isA = (k == A || t == A);
    const int other = isA ? (k == A ? t : k) : (k == B ? t : k);
    return rows[((isA ? 0 : 2) + (logm ? 1 : 0)) * c.cap + other];
  }
  return (logm ? c.WL : c.WD)[tri(k, t, c.cap)];
}

__device__ double loglik_eval(const Ctx& c, const int* sz) {
  const rc_kparams& kp = *c.kp;
  const rc_params& P = kp.P;
  __syncthreads();
  if (threadIdx.x < 32) {                                                    // C = findall(clustsizes .> 0)  (:22)
    int base = 0;
    for (int w = 0; w * 32 < c.cap; ++w) {
      const int s = w * 32 + threadIdx.x;
      const bool live = s < c.cap && sz[s] > 0;
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) c.clist[base + __popc(m & ((1u << threadIdx.x) - 1u))] = (uint8_t)s;
      base += __popc(m);
    }
    if (threadIdx.x == 0) c.sc->itmp[0] = base;
  }
  __syncthreads();
  const int K = c.sc->itmp[0];
  for (int idx = threadIdx.x; idx < K * K; idx += RC_NTHR) {
    const int ki = idx / K, ti = idx - ki * K;
    if (ti < ki) continue;
    const int k = c.clist[ki], t = c.clist[ti];
    const double msD = rc_deq128(getW(c, false, k, t), c.qD);
    const double msL = rc_deq128(getW(c, true, k, t), c.qL);
    double term;
    if (ki == ti) {                                                          // :26-36
      const long long szk = sz[k];
      const double pairs = (double)(szk * (szk - 1) / 2);
      const double a = P.alpha + P.delta1 * pairs;
      const double b = P.beta + msD / 2;
      term = (P.delta1 - 1) * msL / 2 - pairs * kp.lgd1 + kp.abratio + rc_lgamma(a) - a * rc_log(b);
    } else {                                                                 // :40-53
      const double pairs = (double)((long long)sz[k] * (long long)sz[t]);
      const double z = P.zeta + P.delta2 * pairs;
      const double g = P.gamma + msD;
      term = (P.delta2 - 1) * msL - pairs * kp.lgd2 + kp.zgratio + rc_lgamma(z) - z * rc_log(g);
    }
    c.terms[idx] = term;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double L1 = 0;
    for (int ki = 0; ki < K; ++ki) L1 += c.terms[ki * K + ki];
    double L2 = 0;
    for (int ki = 0; ki < K; ++ki)
      for (int ti = ki + 1; ti < K; ++ti) L2 += c.terms[ki * K + ti];
    c.sc->dtmp[0] = P.repulsion ? (L1 + L2) : (L1 + copysign(0.0, L2));      // :54
  }
  __syncthreads();
  return c.sc->dtmp[0];
}

__device__ __forceinline__ double xlogy(double a, double b) { return (a == 0.0 && !rc_isnan(b)) ? 0.0 : a * rc_log(b); }
__device__ __forceinline__ double xlog1py(double a, double b) { return (a == 0.0 && !rc_isnan(b)) ? 0.0 : a * rc_log1p(b); }

// logprior (mcmc.jl:58-78); warp 0.
__device__ double logprior_eval(const Ctx& c) {
  const rc_params& P = c.kp->P;
  const int lane = threadIdx.x & 31;
  const double r = c.sc->r, p = c.sc->p;
  double* tv = c.terms;   // scratch: per-slot terms
  for (int s = lane; s < c.cap; s += 32)
    if (c.sizes[s] > 0) { const double nj = (double)c.sizes[s]; tv[s] = rc_log(nj) + rc_lgamma(nj + r - 1); }
  __syncwarp();
  double L = 0.0;
  if (lane == 0) {
    const double K = (double)c.sc->K;
    const double theta = 1 / P.sigma;
    const double xt = (r > 0 ? r : 0.0) / theta;
    double gl = -rc_lgamma(P.eta) - rc_log(theta) - xt;
    if (xt < RC_INF && xt > -RC_INF) gl += xlogy(P.eta - 1, xt);
    if (r < 0) gl = -RC_INF;
    const double y = p < 0 ? 0.0 : (p > 1 ? 1.0 : p);
    double bl = xlogy(P.u - 1, y) + xlog1py(P.v - 1, -y) - (rc_lgamma(P.u) + rc_lgamma(P.v) - rc_lgamma(P.u + P.v));
    if (p < 0 || p > 1) bl = -RC_INF;
    L = rc_lgamma(K + 1) + ((double)c.n - K) * rc_log(p) + (r * K) * rc_log(1 - p) - K * rc_lgamma(r) + gl + bl;
    for (int s = 0; s < c.cap; ++s)
      if (c.sizes[s] > 0) L += tv[s];
  }
  return __shfl_sync(0xffffffffu, L, 0);
}

// sample_r! (mcmc.jl:80-136); warp 0.  Returns accept.
__device__ bool update_r(const Ctx& c, unsigned it) {
  const rc_params& P = c.kp->P;
  const int lane = threadIdx.x & 31;
  const double r = c.sc->r, p = c.sc->p, sd = P.proposalsd_r;
  double cand = r;
  if (lane == 0) {
    const double lb = (0.0 - r) / sd;
    double z = 0;
    for (uint32_t att = 0; att < 100000u; ++att) {
      z = rc_norminv(rc_open01(rc_draw1(c.key, it, RC_SITE_R_NORMAL, 0, att, 0)));
      if (z >= lb) break;
    }
    cand = r + sd * z;
  }
  cand = __shfl_sync(0xffffffffu, cand, 0);
  double* tc = c.terms;            // per-slot lgamma terms (candidate / current)
  double* tr = c.terms + c.cap;
  for (int s = lane; s < c.cap; s += 32)
    if (c.sizes[s] > 0) {
      const double nk1 = (double)(c.sizes[s] - 1);
      tc[s] = rc_lgamma(nk1 + cand);
      tr[s] = rc_lgamma(nk1 + r);
    }
  __syncwarp();
  int accept = 0;
  if (lane == 0) {
    const double K = (double)c.sc->K;
    const double l1mp = rc_log(1 - p);
    double lpc = (P.eta - 1) * rc_log(cand) + K * (cand * l1mp - rc_lgamma(cand)) - cand * P.sigma;
    double lpr = (P.eta - 1) * rc_log(r) + K * (r * l1mp - rc_lgamma(r)) - r * P.sigma;
    for (int s = 0; s < c.cap; ++s)
      if (c.sizes[s] > 0) { lpc = lpc + tc[s]; lpr = lpr + tr[s]; }
    const double log2pi = 1.8378770664093454836;
    // logpdf(truncated(Normal(mu, sd), 0, Inf), x)
    double lq[2];
    for (int q = 0; q < 2; ++q) {
      const double mu = q == 0 ? r : cand, x = q == 0 ? cand : r;
      const double zz = (x - mu) / sd;
      const double lcdf = rc_normcdf((0.0 - mu) / sd);
      const double logtp = rc_log(1.0 - lcdf);
      lq[q] = -(zz * zz + log2pi) / 2 - rc_log(sd) - logtp;
    }
    const double lratio = lq[0] - lq[1];
    const double lu = rc_log(rc_draw1(c.key, it, RC_SITE_R_ACCEPT, 0, 0, 0));
    accept = lu < rc_min0(lpc - lpr - lratio);
    if (accept) c.sc->r = cand;
  }
  return __shfl_sync(0xffffffffu, accept, 0) != 0;
}

// sample_p! (mcmc.jl:138-155); thread 0.
__device__ void update_p(const Ctx& c, unsigned it) {
  const rc_params& P = c.kp->P;
  const double a = (double)(c.n - c.sc->K) + P.u;
  const double b = c.sc->r * (double)c.sc->K + P.v;
  const double p = rc_beta(a, b, c.key, it);
  c.sc->p = p;
  c.sc->logp = rc_log(p);
  c.sc->log1mp = rc_log(1 - p);
}

// ------------------------------------------------------------------------------------------------
// Restricted Gibbs scan (mcmc.jl:259-354) over Slist[0..nS) on the launch state (labL, szL).
// forced: allocate toward the chain's current labels (c.lab) and only accumulate the probability.
// The log transition probability is returned in c.sc->ltp.
// ------------------------------------------------------------------------------------------------
__device__ void restricted_scan(const Ctx& c, unsigned it, unsigned mh, unsigned scan, int nS, int pi, int pj, int ca,
                                int cb, int c1, int c2, bool forced) {
  const rc_kparams& kp = *c.kp;
  const rc_params& P = kp.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) c.sc->ltp = 0.0;
  for (int pos = 0; pos < nS; ++pos) {
    const int y = c.Slist[pos];
    __syncthreads();
    if (tid == 0) {                                                          // :303-304
      c.szL[c.labL[y]] -= 1;
      c.labL[y] = RC_DETACHED;
    }
    __syncthreads();
    // sums of row y over the current members of the two candidate clusters (all live in S u {i, j})
    long long aD = 0, aL = 0, bD = 0, bL = 0;
    const longlong2* row = c.DL + (size_t)y * c.n;
    for (int q = tid; q < nS + 2; q += RC_NTHR) {
      const int x = q < nS ? (int)c.Slist[q] : (q == nS ? pi : pj);
      const int l = c.labL[x];
      if (l == ca) { const longlong2 v = __ldg(row + x); aD += v.x; aL += v.y; }
      else if (l == cb) { const longlong2 v = __ldg(row + x); bD += v.x; bL += v.y; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      aD += shfl_xor_ll(aD, off); aL += shfl_xor_ll(aL, off);
      bD += shfl_xor_ll(bD, off); bL += shfl_xor_ll(bL, off);
    }
    if (lane == 0) { c.red[warp * 4 + 0] = aD; c.red[warp * 4 + 1] = aL; c.red[warp * 4 + 2] = bD; c.red[warp * 4 + 3] = bL; }
    __syncthreads();
    if (warp == 0) {
      // lanes 0..3 evaluate slots {ca, cb, c1, c2}
      const int slot = lane == 0 ? ca : (lane == 1 ? cb : (lane == 2 ? c1 : c2));
      double L1 = 0.0, lpr = 0.0, L2p = 0.0;
      if (lane < 4) {
        long long sd, sl;
        if (slot == ca || slot == cb) {
          const int o = slot == ca ? 0 : 2;
          sd = 0; sl = 0;
          for (int w = 0; w < RC_NWARP; ++w) { sd += c.red[w * 4 + o]; sl += c.red[w * 4 + o + 1]; }
        } else {
          const longlong2 t = c.T[(size_t)pos * c.cap + slot];
          sd = t.x; sl = t.y;
        }
        const int szs = c.szL[slot];
        const double szd = (double)szs;
        const double sD = rc_dequant(sd, c.qD), sL = rc_dequant(sl, c.qL);
        const double z_i = P.zeta + P.delta2 * szd, g_i = P.gamma + sD;                       // :313-319
        L2p = kp.LGZ[szs] - z_i * rc_log(g_i) + kp.zgratio + (P.delta2 - 1) * sL - szd * kp.lgd2;   // :327-330
        if (lane < 2) {                                                                       // :307-312, 321-326
          const double a_i = P.alpha + P.delta1 * szd, b_i = P.beta + sD;
          L1 = kp.LGA[szs] + kp.abratio - a_i * rc_log(b_i) + (P.delta1 - 1) * sL - szd * kp.lgd1;
          lpr = kp.LOGN[szs + 1] + c.sc->logp + rc_log((double)(szs - 1) + c.sc->r) - kp.LOGN[szs];
        }
      }
      const double L2i = __shfl_sync(0xffffffffu, L2p, 2) + __shfl_sync(0xffffffffu, L2p, 3);   // :331 (quirk Q2)
      const double L2 = L2i - L2p;                                                               // :332-334
      const double lpv = lpr + (L1 + (P.repulsion ? L2 : copysign(0.0, L2)));                    // :335
      double lp0 = __shfl_sync(0xffffffffu, lpv, 0), lp1 = __shfl_sync(0xffffffffu, lpv, 1);
      if (lane == 0) {
        int k, cnew;
        if (!forced) {                                                                           // :336-338
          const rc_draw dr = rc_draw2(c.key, it, RC_SITE_SM_RGIBBS, mh, scan, (uint32_t)pos);
          double mn = lp0;
          if (!rc_isnan(mn)) { if (rc_isnan(lp1) || lp1 < mn) mn = lp1; }
          lp0 -= mn; lp1 -= mn;
          const double g0 = -rc_log(-rc_log(dr.u0)) + lp0;
          const double g1 = -rc_log(-rc_log(dr.u1)) + lp1;
          k = 0;
          if (!rc_isnan(g0)) { if (rc_isnan(g1) || g1 > g0) k = 1; }
          cnew = k == 0 ? ca : cb;
        } else {                                                                                 // :339-342
          cnew = c.lab[y];
          k = (ca == cnew) ? 0 : 1;
        }
        c.labL[y] = (uint8_t)cnew;                                                               // :344-345
        c.szL[cnew] += 1;
        double mn = lp0;                                                                         // :348 (quirk Q3)
        if (!rc_isnan(mn)) { if (rc_isnan(lp1) || lp1 < mn) mn = lp1; }
        lp0 += mn; lp1 += mn;
        double p0 = rc_exp(lp0), p1 = rc_exp(lp1);
        const double den = p0 + p1;
        p0 /= den; p1 /= den;
        c.sc->ltp += rc_log(k == 0 ? p0 : p1);                                                   // :351
      }
    }
  }
  __syncthreads();
}

// One split-merge proposal (mcmc.jl:372-474) on the chain's current state.  Returns accept through
// c.sc->itmp[1], split through itmp[2].
__device__ void splitmerge_step(const Ctx& c, unsigned it, unsigned mh) {
  const rc_kparams& kp = *c.kp;
  const rc_params& P = kp.P;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = c.n, cap = c.cap;
  const double r = c.sc->r, p = c.sc->p;
  const int K = c.sc->K;
  // (i, j) = sample(1:n, 2, replace = false)  (:379)
  const rc_draw dr = rc_draw2(c.key, it, RC_SITE_SM_PAIR, mh, 0, 0);
  long long i1 = rc_randint(dr.u0, n), i2 = rc_randint(dr.u1, n - 1);
  if (i2 == i1) i2 = n;
  const int pi = (int)i1 - 1, pj = (int)i2 - 1;
  const int ci = c.lab[pi], cj = c.lab[pj];
  __syncthreads();
  if (tid == 0) { c.sc->itmp[1] = 0; c.sc->itmp[2] = 0; }
  if (P.maxK > 0 && ci == cj && K >= P.maxK) { __syncthreads(); return; }   // :384-386
  // S = members of ci or cj except i, j, ascending (:389-390): ordered compaction
  {
    const int chunk = (n + RC_NTHR - 1) / RC_NTHR;
    const int b = tid * chunk, e = min(n, b + chunk);
    int cntm = 0;
    for (int k = b; k < e; ++k) cntm += ((c.lab[k] == ci || c.lab[k] == cj) && k != pi && k != pj);
    int incl = cntm;
    for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    if (lane == 31) c.itmp[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += c.itmp[w];
    int o = woff + incl - cntm;
    for (int k = b; k < e; ++k)
      if ((c.lab[k] == ci || c.lab[k] == cj) && k != pi && k != pj) c.Slist[o++] = (unsigned short)k;
    if (tid == RC_NTHR - 1) c.sc->itmp[3] = woff + incl;
    __syncthreads();
  }
  const int nS = c.sc->itmp[3];
  // launch state (:393-408)
  for (int k = tid; k < n; k += RC_NTHR) c.labL[k] = c.lab[k];
  for (int s = tid; s < cap; s += RC_NTHR) c.szL[s] = c.sizes[s];
  __syncthreads();
  const bool split = ci == cj;
  int ca = ci;
  if (split) {
    if (tid < 32) {
      int e = -1;
      for (int w = 0; w * 32 < cap && e < 0; ++w) {
        const int s = w * 32 + lane;
        const unsigned m = __ballot_sync(0xffffffffu, s < cap && c.sizes[s] == 0);
        if (m) e = w * 32 + __ffs(m) - 1;
      }
      if (lane == 0) c.sc->itmp[4] = e;
    }
    __syncthreads();
    ca = c.sc->itmp[4];
    if (ca < 0) { if (tid == 0) c.sc->status = RC_ERR_SLOTS; __syncthreads(); return; }
    if (tid == 0) { c.labL[pi] = (uint8_t)ca; c.szL[ci] -= 1; c.szL[ca] += 1; }
  }
  const int cb = cj;
  __syncthreads();
  {
    int na = 0, nb = 0;   // launch allocation of S (:402-407)
    for (int pos = tid; pos < nS; pos += RC_NTHR) {
      const int k = c.Slist[pos];
      const double u = rc_draw1(c.key, it, RC_SITE_SM_LAUNCH, mh, (uint32_t)pos, 0);
      const int cn = rc_randint(u, 2) == 1 ? ca : cb;
      c.labL[k] = (uint8_t)cn;
      na += cn == ca; nb += cn == cb;
    }
    for (int off = 16; off; off >>= 1) { na += __shfl_xor_sync(0xffffffffu, na, off); nb += __shfl_xor_sync(0xffffffffu, nb, off); }
    if (lane == 0) { c.itmp[warp * 2] = na; c.itmp[warp * 2 + 1] = nb; }
    __syncthreads();
    if (tid == 0) {
      int ta = 0, tb = 0;
      for (int w = 0; w < RC_NWARP; ++w) { ta += c.itmp[w * 2]; tb += c.itmp[w * 2 + 1]; }
      c.szL[ca] = 1 + ta;
      c.szL[cb] = 1 + tb;
    }
    __syncthreads();
  }
  // first two live slots of the launch state (C[1], C[2] of :273, :331)
  if (tid == 0) {
    int c1 = -1, c2 = -1;
    for (int s = 0; s < cap && c2 < 0; ++s)
      if (c.szL[s] > 0) { if (c1 < 0) c1 = s; else c2 = s; }
    c.sc->itmp[5] = c1; c.sc->itmp[6] = c2;
  }
  // row sums by slot of every member of S u {i, j} under the launch labels
  build_perm(c, c.labL);
  for (int pos = 0; pos < nS + 2; ++pos) {
    const int x = pos < nS ? (int)c.Slist[pos] : (pos == nS ? pi : pj);
    reduce_row(c, x);
    __syncthreads();
    for (int s = tid; s < cap; s += RC_NTHR) c.T[(size_t)pos * cap + s] = bin_total(c, s);
    __syncthreads();
  }
  const int c1 = c.sc->itmp[5], c2 = c.sc->itmp[6];
  for (unsigned g = 0; g < (unsigned)kp.numGibbs; ++g)                       // :411-414
    restricted_scan(c, it, mh, g, nS, pi, pj, ca, cb, c1, c2, false);
  double log_prior_ratio, log_proposal_ratio;
  if (split) {                                                              // :416-434
    restricted_scan(c, it, mh, (unsigned)kp.numGibbs, nS, pi, pj, ca, cb, c1, c2, false);
    const int sza = c.szL[ca], szb = c.szL[cb];   // szfinal[cfinal[i]], szfinal[cfinal[j]]
    log_prior_ratio = rc_log((double)(K + 1)) + r * rc_log(1 - p) - rc_log(p) - rc_lgamma(r) +
                      rc_lgamma((double)(sza - 1) + r) + rc_lgamma((double)(szb - 1) + r) +
                      rc_log((double)sza) + rc_log((double)szb) +
                      -(rc_lgamma((double)(c.sizes[ci] - 1) + r) + rc_log((double)c.sizes[ci]));
    log_proposal_ratio = c.sc->ltp;
    // block sums of the proposed state: rows a (= new slot ca) and b (= cb)
    rc_i128* rows = reinterpret_cast<rc_i128*>(c.partial);   // [rowA_D | rowA_L | rowB_D | rowB_L] x cap
    __syncthreads();
    for (int t = tid; t < cap; t += RC_NTHR) {
      rc_i128 sd, sl; sd.lo = 0; sd.hi = 0; sl.lo = 0; sl.hi = 0;
      for (int q = 0; q < nS + 2; ++q) {
        const int x = q < nS ? (int)c.Slist[q] : (q == nS ? pi : pj);
        if (c.labL[x] == ca) { const longlong2 v = c.T[(size_t)q * cap + t]; rc_add128(sd, v.x); rc_add128(sl, v.y); }
      }
      rows[0 * cap + t] = sd; rows[1 * cap + t] = sl;
    }
    __syncthreads();
    // cross = sum_{x in a_F, y in b_F} DL[x][y]: one warp per row x, lanes over the members
    rc_i128 crD, crL; crD.lo = 0; crD.hi = 0; crL.lo = 0; crL.hi = 0;
    for (int q = warp; q < nS + 2; q += RC_NWARP) {
      const int x = q < nS ? (int)c.Slist[q] : (q == nS ? pi : pj);
      if (c.labL[x] != ca) continue;
      const longlong2* row = c.DL + (size_t)x * n;
      long long sd = 0, sl = 0;
      for (int q2 = lane; q2 < nS + 2; q2 += 32) {
        const int y = q2 < nS ? (int)c.Slist[q2] : (q2 == nS ? pi : pj);
        if (c.labL[y] == cb) { const longlong2 v = __ldg(row + y); sd += v.x; sl += v.y; }
      }
      for (int off = 16; off; off >>= 1) { sd += shfl_xor_ll(sd, off); sl += shfl_xor_ll(sl, off); }
      rc_add128(crD, sd); rc_add128(crL, sl);
    }
    rc_i128* red128 = reinterpret_cast<rc_i128*>(c.terms);    // scratch
    if (lane == 0) { red128[warp * 2] = crD; red128[warp * 2 + 1] = crL; }
    __syncthreads();
    if (tid == 0) {
      rc_i128 xD = red128[0], xL = red128[1];
      for (int w = 1; w < RC_NWARP; ++w) { rc_add128(xD, red128[w * 2]); rc_add128(xL, red128[w * 2 + 1]); }
      // sum_{x in a_F} R[x], R[x] = launch-state row sum over all of ci: columns ca + cb of rowA
      rc_i128 totD = rows[0 * cap + ca], totL = rows[1 * cap + ca];
      rc_add128(totD, rows[0 * cap + cb]); rc_add128(totL, rows[1 * cap + cb]);
      rc_i128 aaD = totD, aaL = totL;
      rc_sub128(aaD, xD); rc_sub128(aaL, xL);
      rc_i128 bbD = c.WD[tri(ci, ci, cap)], bbL = c.WL[tri(ci, ci, cap)];
      rc_sub128(bbD, aaD); rc_sub128(bbD, xD); rc_sub128(bbD, xD);
      rc_sub128(bbL, aaL); rc_sub128(bbL, xL); rc_sub128(bbL, xL);
      c.sc->aaD = aaD; c.sc->aaL = aaL; c.sc->abD = xD; c.sc->abL = xL; c.sc->bbD = bbD; c.sc->bbL = bbL;
    }
    __syncthreads();
    for (int t = tid; t < cap; t += RC_NTHR) {                               // row b = row ci of the current state - row a
      rc_i128 bD = c.WD[tri(ci, t, cap)], bL = c.WL[tri(ci, t, cap)];
      rc_sub128(bD, rows[0 * cap + t]); rc_sub128(bL, rows[1 * cap + t]);
      rows[2 * cap + t] = bD; rows[3 * cap + t] = bL;
    }
    if (tid == 0) { c.sc->fslotA = ca; c.sc->fslotB = cb; c.sc->itmp[2] = 1; }
    __syncthreads();
  } else {                                                                  // merge (:435-459)
    const int szf = c.sizes[ci] + c.sizes[cj];
    log_prior_ratio = -(rc_log((double)K) + r * rc_log(1 - p) - rc_log(p) - rc_lgamma(r)) +
                      rc_lgamma((double)(szf - 1) + r) + rc_log((double)szf) +
                      -(rc_lgamma((double)(c.sizes[ci] - 1) + r) + rc_lgamma((double)(c.sizes[cj] - 1) + r) +
                        rc_log((double)c.sizes[ci]) + rc_log((double)c.sizes[cj]));
    restricted_scan(c, it, mh, (unsigned)kp.numGibbs, nS, pi, pj, ca, cb, c1, c2, true);   // :454-455
    log_proposal_ratio = -c.sc->ltp;
    rc_i128* rows = reinterpret_cast<rc_i128*>(c.partial);
    __syncthreads();
    for (int t = tid; t < cap; t += RC_NTHR) {                               // row cj of the merged state
      rc_i128 bD = c.WD[tri(ci, t, cap)], bL = c.WL[tri(ci, t, cap)];
      rc_add128(bD, c.WD[tri(cj, t, cap)]); rc_add128(bL, c.WL[tri(cj, t, cap)]);
      rows[2 * cap + t] = bD; rows[3 * cap + t] = bL;
      rc_i128 z; z.lo = 0; z.hi = 0;
      rows[0 * cap + t] = z; rows[1 * cap + t] = z;
    }
    if (tid == 0) {
      rc_i128 bbD = c.WD[tri(ci, ci, cap)], bbL = c.WL[tri(ci, ci, cap)];
      rc_add128(bbD, c.WD[tri(cj, cj, cap)]); rc_add128(bbL, c.WL[tri(cj, cj, cap)]);
      rc_add128(bbD, c.WD[tri(ci, cj, cap)]); rc_add128(bbD, c.WD[tri(ci, cj, cap)]);
      rc_add128(bbL, c.WL[tri(ci, cj, cap)]); rc_add128(bbL, c.WL[tri(ci, cj, cap)]);
      rc_i128 z; z.lo = 0; z.hi = 0;
      c.sc->aaD = z; c.sc->aaL = z; c.sc->abD = z; c.sc->abL = z; c.sc->bbD = bbD; c.sc->bbL = bbL;
      c.sc->fslotA = ci; c.sc->fslotB = cj;
    }
    __syncthreads();
    // sizes of the merged state
    for (int s = tid; s < cap; s += RC_NTHR) c.szL[s] = c.sizes[s];
    __syncthreads();
    if (tid == 0) { c.szL[ci] = 0; c.szL[cj] = szf; }
    __syncthreads();
  }
  const double ll_fin = loglik_eval(c, c.szL);                              // :462-464
  __syncthreads();
  if (tid == 0) { c.sc->fslotA = -1; c.sc->fslotB = -1; }
  __syncthreads();
  const double ll_cur = loglik_eval(c, c.sizes);
  if (tid == 0) {
    const double log_lik_ratio = ll_fin - ll_cur;
    const double lar = rc_min0(log_prior_ratio + log_lik_ratio - log_proposal_ratio);   // :467-468
    const double lu = rc_log(rc_draw1(c.key, it, RC_SITE_SM_ACCEPT, mh, 0, 0));
    c.sc->itmp[1] = lu < lar ? 1 : 0;                                                   // :469-472
  }
  __syncthreads();
}

// sortlabels (utils.jl:69-74): first-appearance relabelling to 1..K.
__device__ void record_labels(const Ctx& c, uint8_t* out) {
  const int tid = threadIdx.x;
  for (int s = tid; s < c.cap; s += RC_NTHR) c.itmp[s] = 0x7fffffff;
  __syncthreads();
  for (int j = tid; j < c.n; j += RC_NTHR) atomicMin(&c.itmp[c.lab[j]], j);
  __syncthreads();
  for (int s = tid; s < c.cap; s += RC_NTHR) {
    const int f = c.itmp[s];
    int id = 1;
    for (int t = 0; t < c.cap; ++t) id += c.itmp[t] < f;
    c.clist[s] = (uint8_t)id;
  }
  __syncthreads();
  for (int j = tid; j < c.n; j += RC_NTHR) out[j] = c.clist[c.lab[j]];
  __syncthreads();
}

__global__ void __launch_bounds__(RC_NTHR) k_chain(const __grid_constant__ rc_kparams kp) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int chain = blockIdx.x;
  const int tid = threadIdx.x;
  const int n = kp.n, cap = kp.cap, tiles = kp.tiles;
  Ctx c;
  c.n = n; c.cap = cap; c.tiles = tiles; c.qD = kp.qD; c.qL = kp.qL; c.DL = kp.DL; c.kp = &kp;
  {
    size_t o = 0;
    auto take = [&](size_t bytes) { unsigned char* p = smem + o; o += (bytes + 15) & ~(size_t)15; return p; };
    c.partial = reinterpret_cast<longlong2*>(take(sizeof(longlong2) * RC_NWARP * cap));
    c.sc = reinterpret_cast<Scal*>(take(sizeof(Scal)));
    c.red = reinterpret_cast<long long*>(take(sizeof(long long) * RC_NWARP * 4));
    c.perm = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * kp.npad_max));
    c.runStart = reinterpret_cast<unsigned short*>(take(sizeof(unsigned short) * (tiles * cap + 1)));
    c.cnt = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * tiles * cap));
    c.tileStart = reinterpret_cast<int*>(take(sizeof(int) * (tiles + 1)));
    c.sizes = reinterpret_cast<int*>(take(sizeof(int) * cap));
    c.szL = reinterpret_cast<int*>(take(sizeof(int) * cap));
    c.itmp = reinterpret_cast<int*>(take(sizeof(int) * cap));
    c.clist = reinterpret_cast<uint8_t*>(take(cap));
    c.glabel = reinterpret_cast<uint8_t*>(take(kp.npad_max / RC_GROUP));
    c.lab = reinterpret_cast<uint8_t*>(take(n));
    c.labL = reinterpret_cast<uint8_t*>(take(n));
  }
  c.WD = kp.WD + (size_t)chain * cap * cap;
  c.WL = kp.WL + (size_t)chain * cap * cap;
  c.T = kp.T + (size_t)chain * n * cap;
  c.Slist = kp.Slist + (size_t)chain * n;
  c.terms = kp.terms + (size_t)chain * cap * cap;
  c.key = rc_chain_key(kp.seed, (unsigned long long)(kp.chain_offset + chain));

  // load the chain's state
  for (int j = tid; j < n; j += RC_NTHR) c.lab[j] = kp.labels[(size_t)chain * n + j];
  for (int s = tid; s < cap; s += RC_NTHR) c.sizes[s] = kp.sizes[(size_t)chain * cap + s];
  if (tid == 0) {
    Scal& s = *c.sc;
    s.r = kp.r[chain]; s.p = kp.p[chain];
    s.logp = rc_log(s.p); s.log1mp = rc_log(1 - s.p);
    s.status = kp.status[chain]; s.rebuild = 0; s.fslotA = -1; s.fslotB = -1; s.ltp = 0.0;
    int K = 0;
    for (int q = 0; q < cap; ++q) K += kp.sizes[(size_t)chain * cap + q] > 0;
    s.K = K;
  }
  __syncthreads();
  build_perm(c, c.lab);
  if (kp.init_W) init_W(c);
  if (kp.loglik_only) {
    const double ll = loglik_eval(c, c.sizes);
    if (tid == 0) kp.out_ll[chain] = ll;
    return;
  }

  for (long long iter = kp.it0 + 1; iter <= kp.it1; ++iter) {
    if (c.sc->status) break;
    const unsigned it = (unsigned)iter;
    if (tid < 32) {
      const bool ra = update_r(c, it);                                       // mcmc.jl:538
      if (tid == 0) {
        kp.r_acc[(size_t)chain * kp.numiters + (iter - 1)] = ra ? 1 : 0;
        update_p(c, it);                                                     // :539
      }
    }
    __syncthreads();
    // sample_labels! (:540)
    bool accepted_any = false;
    bool perm_dirty = false;
    for (unsigned mh = 0; mh < (unsigned)kp.numMH; ++mh) {
      splitmerge_step(c, it, mh);
      if (c.sc->status) break;
      const int acc = c.sc->itmp[1], spl = c.sc->itmp[2];
      if (tid == 0) {
        kp.sm_acc[((size_t)chain * kp.numiters + (iter - 1)) * kp.numMH + mh] = (uint8_t)acc;
        kp.sm_split[((size_t)chain * kp.numiters + (iter - 1)) * kp.numMH + mh] = (uint8_t)spl;
      }
      perm_dirty = true;
      if (acc) { accepted_any = true; break; }   // numMH == 1: the accepted state never reaches the caller (quirk Q1)
      __syncthreads();
    }
    __syncthreads();
    if (c.sc->status) break;
    if (perm_dirty) build_perm(c, c.lab);
    // Quirk Q1: an accepted proposal rebinds sample_labels!'s LOCAL state; the final scan
    // (:477) then runs on that local object and the caller's labels are untouched this iteration.  The
    // draws of that discarded scan are independent of everything kept (structured stream), so it is skipped.
    if (!accepted_any) full_scan(c, it);
    if (c.sc->status) break;
    if (iter > kp.burnin && (iter - kp.burnin) % kp.thin == 0) {             // :546-554
      const long long j = (iter - kp.burnin) / kp.thin - 1;
      if (j < kp.numsamples) {
        record_labels(c, kp.out_labels + ((size_t)chain * kp.numsamples + j) * n);
        const double ll = loglik_eval(c, c.sizes);
        if (tid < 32) {
          const double lpv = logprior_eval(c);
          if (tid == 0) {
            const size_t o = (size_t)chain * kp.numsamples + j;
            kp.out_K[o] = c.sc->K; kp.out_r[o] = c.sc->r; kp.out_p[o] = c.sc->p;
            kp.out_ll[o] = ll; kp.out_lp[o] = ll + lpv;
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  // store the chain's state
  for (int j = tid; j < n; j += RC_NTHR) kp.labels[(size_t)chain * n + j] = c.lab[j];
  for (int s = tid; s < cap; s += RC_NTHR) kp.sizes[(size_t)chain * cap + s] = c.sizes[s];
  if (tid == 0) { kp.r[chain] = c.sc->r; kp.p[chain] = c.sc->p; kp.status[chain] = c.sc->status; }
}

__global__ void k_tables(rc_params P, int n, double* LGA, double* LGZ, double* LOGN) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= n + 1; s += gridDim.x * blockDim.x) {
    const double sd = (double)s;
    LGA[s] = rc_lgamma(P.alpha + P.delta1 * sd);
    LGZ[s] = rc_lgamma(P.zeta + P.delta2 * sd);
    LOGN[s] = rc_log(sd);
  }
}

}  // namespace

size_t rc_sampler_smem_bytes(int n, int cap, int tiles, int npad_max) {
  size_t o = 0;
  auto take = [&](size_t bytes) { o += (bytes + 15) & ~(size_t)15; };
  take(sizeof(longlong2) * RC_NWARP * cap);
  take(sizeof(Scal));
  take(sizeof(long long) * RC_NWARP * 4);
  take(sizeof(unsigned short) * npad_max);
  take(sizeof(unsigned short) * (tiles * cap + 1));
  take(sizeof(unsigned int) * tiles * cap);
  take(sizeof(int) * (tiles + 1));
  take(sizeof(int) * cap);
  take(sizeof(int) * cap);
  take(sizeof(int) * cap);
  take(cap);
  take(npad_max / RC_GROUP);
  take(n);
  take(n);
  return o;
}

void rc_launch_chain_kernel(const rc_kparams& kp, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_chain<<<kp.nchains, RC_NTHR, smem, st>>>(kp);
}

void rc_launch_tables(const rc_params& P, int n, double* LGA, double* LGZ, double* LOGN, cudaStream_t st) {
  k_tables<<<(n + 2 + 255) / 256, 256, 0, st>>>(P, n, LGA, LGZ, LOGN);
}
