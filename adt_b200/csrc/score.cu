// K7: full-catalog scoring with fused per-user top-K.
// Replaces `predict(full=True)` + host-side masking/argpartition/sort of the whole score row
// (/root/reference/sasrec/model.py:91-96, sasrec/utils.py:718-731, stosa/trainer.py:604-614).
//
// Exact path (this file, fp32 FFMA tiles): scores never leave the SM; every CTA keeps a sorted top-K list per
// user in shared memory and only [splits][U][K] (score,id) pairs are written; a second kernel merges the splits.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/adt_b200.h"
#include "common.cuh"

using namespace adt;

namespace {

constexpr int UT = 64;        // users per CTA
constexpr int KMAX = 64;

struct TopkArgs {
  const float* feats; const float* E; const int* seen_indptr; const int* seen_idx;
  float* part_scores; int* part_ids; float* out_scores; int* out_ids;
  int U, H, n_items, item_offset, K, n_splits;
};

// first position p in [lo, hi) of the sorted id list with idx[p] >= item
__device__ __forceinline__ int seen_lower_bound(const int* __restrict__ idx, int lo, int hi, int item) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (idx[mid] < item) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// [sb, se) = the part of the user's sorted seen list that falls into this CTA's catalog split (found once per CTA): for
// most (user, split) pairs it is empty and the test costs no memory access at all
__device__ __forceinline__ bool is_seen(const TopkArgs& a, int sb, int se, int item) {
  if (sb >= se) return false;
  const int p = seen_lower_bound(a.seen_idx, sb, se, item);
  return p < se && a.seen_idx[p] == item;
}

// better(a,b): a ranks strictly before b  (score desc, then id asc -- deterministic under exact ties)
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

__global__ void __launch_bounds__(NT) score_topk_kernel(TopkArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, K = a.K;
  const int ld = H + 4;
  float* Fs = smem;                       // [UT][ld]
  float* Sc = Fs + UT * ld;               // [UT][CHP]
  float* Ws = Sc + UT * CHP;              // staging
  float* Ls = Ws + WS_FLOATS;             // [UT][K] scores
  int* Li = reinterpret_cast<int*>(Ls + UT * K);   // [UT][K] ids
  int* Ln = Li + UT * K;                  // [UT] fill counts
  int* Sb = Ln + UT;                      // [UT] seen-list range of this split
  int* Se = Sb + UT;
  const int split = blockIdx.x, u0 = blockIdx.y * UT;
  const int per = (a.n_items + a.n_splits - 1) / a.n_splits;
  const int it0 = split * per, it1 = min(a.n_items, it0 + per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;

  load_tile<UT>(Fs, ld, a.feats, H, 0, H, u0, a.U);
  for (int i = threadIdx.x; i < UT; i += NT) {
    Ln[i] = 0;
    int sb = 0, se = 0;
    if (a.seen_indptr && u0 + i < a.U) {
      const int hi = a.seen_indptr[u0 + i + 1];
      sb = seen_lower_bound(a.seen_idx, a.seen_indptr[u0 + i], hi, a.item_offset + it0);
      se = seen_lower_bound(a.seen_idx, sb, hi, a.item_offset + it1);
    }
    Sb[i] = sb;
    Se[i] = se;
  }
  __syncthreads();

  const int nrc = (H + CH - 1) / CH;
  for (int c0 = it0; c0 < it1; c0 += CH) {
    const int ncols = min(CH, it1 - c0);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    stage_chunk<true, CHP>(Ws, a.E, H, c0, 0, ncols, min(CH, H));
    cp_async_commit();
    for (int rc = 0; rc < nrc; ++rc) {
      if (rc + 1 < nrc) {
        stage_chunk<true, CHP>(Ws + ((rc + 1) & 1) * WS_SLOT, a.E, H, c0, (rc + 1) * CH, ncols, min(CH, H - (rc + 1) * CH));
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      mma_nt<4>(acc, Fs, ld, rc * CH, Ws + (rc & 1) * WS_SLOT, min(CH, H - rc * CH));
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(Sc + (ty + 16 * i) * CHP + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    // selection: warp per user row.  Lanes filter their two columns against the row's K-th best (and the seen list); the
    // survivors are merged into the sorted list IN PARALLEL by rank counting -- every survivor and every list entry
    // computes its position in the merged order ((score desc, id asc) is a strict total order, so positions are unique)
    // and writes itself there.  Cost per merge ~ (#survivors + K) short steps instead of #survivors serial insertions,
    // which matters because a split of a small catalog never gets a warm list.
    for (int r = w; r < UT; r += NT / 32) {
      const int u = u0 + r;
      if (u >= a.U) continue;
      float* ls = Ls + r * K;
      int* li = Li + r * K;
      const int sb = Sb[r], se = Se[r];
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int col = l + 32 * half;
        const float s = Sc[r * CHP + col];
        const int item = a.item_offset + c0 + col;
        const int n = Ln[r];
        const bool cand = col < ncols && (n < K || better(s, item, ls[K - 1], li[K - 1])) && !is_seen(a, sb, se, item);
        const unsigned ballot = __ballot_sync(0xffffffffu, cand);
        if (ballot == 0u) continue;
        // this lane's list entries (K <= 64: slots l and l + 32), read before anything is overwritten
        const bool h0 = l < n, h1 = l + 32 < n;
        const float e0s = h0 ? ls[l] : 0.f, e1s = h1 ? ls[l + 32] : 0.f;
        const int e0i = h0 ? li[l] : 0, e1i = h1 ? li[l + 32] : 0;
        int pc = 0, p0 = l, p1 = l + 32;      // merged positions of: my survivor, my two list entries
        for (unsigned b = ballot; b; b &= b - 1) {
          const int src = __ffs(b) - 1;
          const float cs = __shfl_sync(0xffffffffu, s, src);
          const int ci = __shfl_sync(0xffffffffu, item, src);
          pc += better(cs, ci, s, item) ? 1 : 0;
          p0 += better(cs, ci, e0s, e0i) ? 1 : 0;
          p1 += better(cs, ci, e1s, e1i) ? 1 : 0;
        }
        if (cand)
          for (int j = 0; j < n; ++j) pc += better(ls[j], li[j], s, item) ? 1 : 0;   // same address in all lanes: broadcast
        __syncwarp();
        if (h0 && p0 < K) { ls[p0] = e0s; li[p0] = e0i; }
        if (h1 && p1 < K) { ls[p1] = e1s; li[p1] = e1i; }
        if (cand && pc < K) { ls[pc] = s; li[pc] = item; }
        if (l == 0) Ln[r] = min(K, n + __popc(ballot));
        __syncwarp();
      }
    }
    __syncthreads();
  }
  // write this split's lists (sorted, padded with -inf / -1)
  for (int i = threadIdx.x; i < UT * K; i += NT) {
    const int r = i / K, k = i - r * K;
    if (u0 + r < a.U) {
      const bool ok = k < Ln[r];
      const long long o = ((long long)split * a.U + (u0 + r)) * K + k;
      a.part_scores[o] = ok ? Ls[r * K + k] : -INFINITY;
      a.part_ids[o] = ok ? Li[r * K + k] : -1;
    }
  }
}

// warp per user: K-way selection over the n_splits sorted partial lists
__global__ void __launch_bounds__(256) topk_merge_kernel(TopkArgs a) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int u = blockIdx.x * 8 + w;
  if (u >= a.U) return;
  const int K = a.K, S = a.n_splits;
  // lane owns splits l, l+32, ... ; head pointer per owned split (S <= 32*8)
  int hp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) hp[j] = 0;
  for (int k = 0; k < K; ++k) {
    float bs = -INFINITY;
    int bi = 0x7fffffff, bj = -1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int sp = l + 32 * j;
      if (sp < S && hp[j] < K) {
        const long long o = ((long long)sp * a.U + u) * K + hp[j];
        const float s = a.part_scores[o];
        const int id = a.part_ids[o];
        if (id >= 0 && better(s, id, bs, bi)) { bs = s; bi = id; bj = j; }
      }
    }
    // warp argmax
    float ws = bs; int wi = bi; int wl = bj >= 0 ? l : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (ol >= 0 && (wl < 0 || better(os, oi, ws, wi))) { ws = os; wi = oi; wl = ol; }
    }
    if (wl == l && bj >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j == bj) hp[j]++;
    }
    if (l == 0) {
      a.out_scores[(long long)u * K + k] = wl >= 0 ? ws : -INFINITY;
      a.out_ids[(long long)u * K + k] = wl >= 0 ? wi : -1;
    }
  }
}


}  // namespace

extern "C" int adt_score_topk(const adt_score_topk_args* a, adt_stream_t s_) {
  cudaStream_t s = (cudaStream_t)s_;
  if (a->K <= 0 || a->K > KMAX || a->H > 256 || (a->H & 3) || a->n_splits <= 0 || a->n_splits > 256 || a->U <= 0) return ADT_E_SHAPE;
  TopkArgs k;
  k.feats = a->feats; k.E = a->item_emb; k.seen_indptr = a->seen_indptr; k.seen_idx = a->seen_idx;
  k.part_scores = a->part_scores; k.part_ids = a->part_ids; k.out_scores = a->out_scores; k.out_ids = a->out_ids;
  k.U = a->U; k.H = a->H; k.n_items = a->n_items; k.item_offset = a->item_offset; k.K = a->K; k.n_splits = a->n_splits;
  const size_t smem = ((size_t)UT * (a->H + 4) + (size_t)UT * CHP + WS_FLOATS + (size_t)UT * a->K * 2 + 3 * UT) * sizeof(float);
  cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(a->n_splits, (a->U + UT - 1) / UT);
  score_topk_kernel<<<grid, NT, smem, s>>>(k);
  if (a->out_ids) topk_merge_kernel<<<(a->U + 7) / 8, 256, 0, s>>>(k);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? ADT_OK : ADT_E_CUDA;
}
