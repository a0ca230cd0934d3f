// Device-side batch assembly and negative / candidate sampling: the CPU DataLoader workers of the reference
// (WarpDataset.sample_data + random_neq, sasrec/utils.py:288-307, :73-77; EvalDataset.sample_data :162-191 with
// PopularSampler.get_negative_samples :57-69) cannot feed a step that takes half a millisecond, so the batches are built on the GPU
// from the user histories kept resident in HBM as CSR.
//
// Randomness is counter based (Philox4x32-10, the generator of the dropout sites): draw k of (user, position) is a pure function of
// (seed, epoch, user, position, k), so a batch does not depend on which other users share it, on the batch size or on the rank, and
// oracle/sampler_oracle.py reproduces every id bit for bit.
//   training negatives : uniform over 1..itemnum, rejected while in the user's history          (random_neq)
//   eval candidates    : popularity-weighted WITHOUT replacement over ids 0..itemnum-1 (the reference's np.random.choice(range(itemnum),
//                        p=popular_p, replace=False): id 0 has p = 0 and id itemnum can never be drawn, quirk B8), skipping seen items:
//                        alias-table draws with rejection of seen / already drawn ids == successive sampling, the same distribution
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adt_b200.h"
#include "common.cuh"

using namespace adt;

namespace {

__device__ __forceinline__ bool sorted_has(const int* __restrict__ s, int lo, int hi, int x) {
  int l = lo, h = hi;
  while (l < h) {
    const int mid = (l + h) >> 1;
    if (__ldg(s + mid) < x) l = mid + 1; else h = mid;
  }
  return l < hi && __ldg(s + l) == x;
}

// uniform integer in [0, n) from a 32-bit word (multiply-shift; bias < n / 2^32)
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t n) { return (uint32_t)(((unsigned long long)r * n) >> 32); }

struct TrainBatchArgs {
  const int* users; const int* indptr; const int* items; const int* sorted_items;
  int* seq; int* dec; int* pos; int* neg;
  int B, L, itemnum; uint32_t seed_lo, seed_hi, epoch;
};

// one thread per (sample, position).  History h[0..n-1] of the user: position idx = L-1-j (j = 0 .. min(n-1, L)-1) holds
// seq = h[n-2-j], pos = h[n-1-j], dec[idx+1] = seq[idx] (shift right by one, utils.py:299-300), neg = uniform not-in-history.
__global__ void __launch_bounds__(256) assemble_train_kernel(TrainBatchArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.L) return;
  const int b = i / a.L, idx = i - b * a.L;
  const int user = a.users[b];
  const int lo = a.indptr[user], hi = a.indptr[user + 1], n = hi - lo;
  const int j = a.L - 1 - idx;
  int s = 0, p = 0, ng = 0;
  if (n >= 2 && j < n - 1) {
    s = a.items[lo + n - 2 - j];
    p = a.items[lo + n - 1 - j];
    if (p != 0) {
      for (uint32_t k = 0;; ++k) {     // random_neq: redraw while the item is in the user's history
        const uint4 r = philox4x32_10((uint32_t)user, (uint32_t)idx, k >> 2, a.epoch, a.seed_lo, a.seed_hi);
        const uint32_t w = (k & 3) == 0 ? r.x : (k & 3) == 1 ? r.y : (k & 3) == 2 ? r.z : r.w;
        ng = 1 + (int)bounded(w, (uint32_t)a.itemnum);
        if (!sorted_has(a.sorted_items, lo, hi, ng) || k > 4096u) break;
      }
    }
  }
  a.seq[i] = s; a.pos[i] = p; a.neg[i] = ng;
  if (idx + 1 < a.L) a.dec[i + 1] = s;
  if (idx == 0) a.dec[i] = 0;
}

struct EvalBatchArgs {
  const int* users; const int* indptr; const int* items; const int* sorted_seen_indptr; const int* sorted_seen;
  const int* last_item;          // [U] optional: item appended at the END of the sequence (test mode: the validation item), 0 = none
  const int* answers;            // [U] held-out item -> column 0 of item_idx
  const float* alias_prob; const int* alias_idx;   // Vose alias table over ids 0..itemnum-1
  int* seq; int* item_idx;
  int U, L, itemnum, S; uint32_t seed_lo, seed_hi, epoch;
};

// one warp per user: the sequence (last L history items, right aligned) and 1 + S candidates
__global__ void __launch_bounds__(256) assemble_eval_kernel(EvalBatchArgs a) {
  extern __shared__ int acc_smem[];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int ui = blockIdx.x * 8 + w;
  if (ui >= a.U) return;
  const int user = a.users[ui];
  const int lo = a.indptr[user], hi = a.indptr[user + 1], n = hi - lo;
  const int extra = a.last_item ? a.last_item[ui] : 0;
  for (int idx = l; idx < a.L; idx += 32) {
    const int j = a.L - 1 - idx;           // 0 = the last position
    int v = 0;
    if (extra != 0) {                      // test mode: the validation item closes the sequence (utils.py:176-183)
      if (j == 0) v = extra;
      else if (j - 1 < n) v = a.items[lo + n - j];
    } else if (j < n) {
      v = a.items[lo + n - 1 - j];
    }
    a.seq[(long long)ui * a.L + idx] = v;
  }
  if (!a.item_idx) return;
  int* out = a.item_idx + (long long)ui * (a.S + 1);
  if (l == 0) out[0] = a.answers[ui];
  const int slo = a.sorted_seen_indptr[user], shi = a.sorted_seen_indptr[user + 1];
  int* mine = acc_smem + w * a.S;     // accepted ids so far
  int cnt = 0;
  for (uint32_t round = 0; cnt < a.S && round < 4096u; ++round) {
    // lane l evaluates draw k = 32*round + l: one Philox call -> (column, coin)
    const uint32_t k = 32u * round + (uint32_t)l;
    const uint4 r = philox4x32_10((uint32_t)user, k, 0x5eedu, a.epoch, a.seed_lo, a.seed_hi);
    const uint32_t col = bounded(r.x, (uint32_t)a.itemnum);
    const float coin = (float)(r.y >> 8) * (1.0f / 16777216.0f);
    const int cand = coin < __ldg(a.alias_prob + col) ? (int)col : __ldg(a.alias_idx + col);
    bool ok = !sorted_has(a.sorted_seen, slo, shi, cand);
    for (int q = 0; ok && q < cnt; ++q) ok = mine[q] != cand;
    // duplicates inside the round: an earlier lane with the same id wins
    for (int src = 0; src < 31; ++src) {
      const int other = __shfl_sync(0xffffffffu, cand, src);
      const bool other_ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, src) != 0;
      if (src < l && other_ok && other == cand) ok = false;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int rank = __popc(bal & ((1u << l) - 1u));
    if (ok && cnt + rank < a.S) { mine[cnt + rank] = cand; out[1 + cnt + rank] = cand; }
    cnt = min(a.S, cnt + __popc(bal));
    __syncwarp();
  }
}

// ---- Bert4Rec cloze instances (bert4rec/datasets/dataset.py:70-158) -------------------------------------------------------------
struct ClozeArgs {
  const int* users; const int* win_start; const int* win_len; const int* dup;     // per instance; dup < 0: the "mask last" instance
  const int* indptr; const int* items;
  int* tokens; int* dec_tokens; int* labels;
  int B, L, itemnum, mask_token; float mask_prob; uint32_t seed_lo, seed_hi, epoch;
};

// one thread per (instance, position).  Window w = history[start, start + len) right aligned in L slots; position j of the window:
//   p = U[0,1) ; p < mask_prob: label = item, token = MASK (p/mask_prob < 0.8) | random item in 1..itemnum (< 0.9) | item ; else token = item
// the decoder copy equals the tokens except that the LAST position is always the mask token (dataset.py:150); the mask-last instance
// masks only the last position (dataset.py:99-121).  Draw = Philox(ctr = (user, start + j, dup, epoch)).
__global__ void __launch_bounds__(256) cloze_kernel(ClozeArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.L) return;
  const int b = i / a.L, idx = i - b * a.L;
  const int user = a.users[b], start = a.win_start[b], len = a.win_len[b], dup = a.dup[b];
  const int j = idx - (a.L - len);          // position inside the window (negative: left padding)
  int tok = 0, dtok = 0, lab = 0;
  if (j >= 0) {
    const int item = a.items[a.indptr[user] + start + j];
    tok = dtok = item;
    if (dup < 0) {
      if (j == len - 1) { tok = dtok = a.mask_token; lab = item; }
    } else {
      const uint4 r = philox4x32_10((uint32_t)user, (uint32_t)(start + j), (uint32_t)dup, a.epoch, a.seed_lo, a.seed_hi);
      float p = (float)(r.x >> 8) * (1.0f / 16777216.0f);
      if (p < a.mask_prob) {
        p = p / a.mask_prob;
        if (p < 0.8f) tok = a.mask_token;
        else if (p < 0.9f) tok = 1 + (int)bounded(r.y, (uint32_t)a.itemnum);
        dtok = tok;
        lab = item;
      }
      if (j == len - 1) dtok = a.mask_token;
    }
  }
  a.tokens[i] = tok; a.dec_tokens[i] = dtok; a.labels[i] = lab;
}

}  // namespace

extern "C" int adt_cloze_batch(const adt_cloze_batch_args* a, adt_stream_t s_) {
  if (a->B <= 0 || a->L <= 0 || a->itemnum <= 0) return ADT_E_SHAPE;
  ClozeArgs k;
  k.users = a->users; k.win_start = a->win_start; k.win_len = a->win_len; k.dup = a->dup; k.indptr = a->hist_indptr; k.items = a->hist_items;
  k.tokens = a->tokens; k.dec_tokens = a->dec_tokens; k.labels = a->labels; k.B = a->B; k.L = a->L; k.itemnum = a->itemnum;
  k.mask_token = a->mask_token; k.mask_prob = a->mask_prob;
  k.seed_lo = (uint32_t)(a->seed & 0xffffffffull); k.seed_hi = (uint32_t)(a->seed >> 32); k.epoch = a->epoch;
  const int n = a->B * a->L;
  cloze_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)s_>>>(k);
  return cudaGetLastError() == cudaSuccess ? ADT_OK : ADT_E_CUDA;
}

extern "C" int adt_assemble_train_batch(const adt_train_batch_args* a, adt_stream_t s_) {
  if (a->B <= 0 || a->L <= 0 || a->itemnum <= 0) return ADT_E_SHAPE;
  TrainBatchArgs k;
  k.users = a->users; k.indptr = a->hist_indptr; k.items = a->hist_items; k.sorted_items = a->hist_sorted;
  k.seq = a->seq; k.dec = a->dec; k.pos = a->pos; k.neg = a->neg; k.B = a->B; k.L = a->L; k.itemnum = a->itemnum;
  k.seed_lo = (uint32_t)(a->seed & 0xffffffffull); k.seed_hi = (uint32_t)(a->seed >> 32); k.epoch = a->epoch;
  const int n = a->B * a->L;
  assemble_train_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)s_>>>(k);
  return cudaGetLastError() == cudaSuccess ? ADT_OK : ADT_E_CUDA;
}

extern "C" int adt_assemble_eval_batch(const adt_eval_batch_args* a, adt_stream_t s_) {
  if (a->U <= 0 || a->L <= 0 || a->itemnum <= 0 || a->n_candidates < 0 || a->n_candidates > 1024) return ADT_E_SHAPE;
  EvalBatchArgs k;
  k.users = a->users; k.indptr = a->hist_indptr; k.items = a->hist_items; k.sorted_seen_indptr = a->seen_indptr; k.sorted_seen = a->seen_sorted;
  k.last_item = a->last_item; k.answers = a->answers; k.alias_prob = a->alias_prob; k.alias_idx = a->alias_idx;
  k.seq = a->seq; k.item_idx = a->n_candidates > 0 ? a->item_idx : nullptr; k.U = a->U; k.L = a->L; k.itemnum = a->itemnum; k.S = a->n_candidates;
  k.seed_lo = (uint32_t)(a->seed & 0xffffffffull); k.seed_hi = (uint32_t)(a->seed >> 32); k.epoch = a->epoch;
  assemble_eval_kernel<<<(a->U + 7) / 8, 256, (size_t)8 * (a->n_candidates > 0 ? a->n_candidates : 1) * sizeof(int), (cudaStream_t)s_>>>(k);
  return cudaGetLastError() == cudaSuccess ? ADT_OK : ADT_E_CUDA;
}
