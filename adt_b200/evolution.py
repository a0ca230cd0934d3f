"""Candidate evaluation of ADT's evolutionary lambda search, one candidate per GPU.

The reference evaluates candidates strictly sequentially (sasrec/evolution.py:172-206: `get_cand_auc` = `_set_choice`
+ `model.eval()` + `evaluate_loader(val)` -> AUC, then GA bookkeeping on the host).  Each evaluation is independent
given the frozen supernet weights, so candidate c goes to rank c mod G; ranks gather the 3 floats per candidate
(AUC, NDCG@10, HR@10) and rank 0 runs the GA bookkeeping unchanged.  The only collective is that all_gather.
"""
import numpy as np
import torch

from .evaluate import rank_of_first_candidate, sampled_metrics, sampled_rank, sampled_metrics_from_acc
from .lambdas import candidate_to_lambdas


def set_choice_from_candidate(model, cand, rec_choice, ind_choice):
    """sasrec/evolution.py:139-153: candidate in [0,1]^(2*layers) -> per-layer lambdas -> model.set_choice."""
    rec, ind = candidate_to_lambdas(list(cand), rec_choice, ind_choice)
    block = []
    for r, i in zip(rec, ind):
        block += [r, i]
    model.set_choice(np.array(block))
    return rec, ind


@torch.no_grad()
def candidate_fitness(model, val_batches, k=10):
    """get_cand_auc (evolution.py:172-179): one validation pass -> (AUC, NDCG@k, HR@k).
    val_batches: iterable of (user_ids, seq [U,L], item_idx [U,C]) with the answer in column 0 (utils.py:162-191)."""
    model.eval()
    ranks, C = [], None
    for u, seq, item_idx in val_batches:
        ranks.append(rank_of_first_candidate(model, u, seq, item_idx).cpu())
        C = item_idx.shape[1] if hasattr(item_idx, "shape") else np.asarray(item_idx).shape[1]
    (ndcg, hr), auc = sampled_metrics(torch.cat(ranks), C, ks=(k,))
    return auc, ndcg[k], hr[k]


def assign(n_candidates, world, rank):
    """indices of the candidates this rank evaluates (round robin: candidate c -> rank c mod world)."""
    return list(range(rank, n_candidates, world))


def evaluate_population(candidates, fitness_fn, process_group=None, local=None):
    """Evaluate `candidates` (list of vectors) in parallel over the ranks of `process_group`.
    fitness_fn(cand) -> tuple of floats is called only for this rank's candidates (or `local` = their already computed tuples).
    Returns an array [n_candidates, n_metrics] identical on every rank."""
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    world = torch.distributed.get_world_size(process_group) if dist_on else 1
    rank = torch.distributed.get_rank(process_group) if dist_on else 0
    mine = assign(len(candidates), world, rank)
    if local is None:
        local = [tuple(float(x) for x in fitness_fn(candidates[c])) for c in mine]
    local = [tuple(float(x) for x in t) for t in local]
    nm = len(local[0]) if local else 0
    if world == 1:
        return np.array(local, dtype=np.float64).reshape(len(candidates), -1)
    backend = torch.distributed.get_backend(process_group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    width = torch.tensor([nm], dtype=torch.int64, device=dev)
    torch.distributed.all_reduce(width, op=torch.distributed.ReduceOp.MAX, group=process_group)
    nm = int(width.item())
    per = (len(candidates) + world - 1) // world
    buf = torch.full((per, nm), float("nan"), dtype=torch.float64, device=dev)
    if local:
        buf[:len(local)] = torch.tensor(local, dtype=torch.float64, device=dev)
    out = torch.empty(world, per, nm, dtype=torch.float64, device=dev)
    torch.distributed.all_gather_into_tensor(out.view(world * per, nm), buf, group=process_group)
    res = np.full((len(candidates), nm), np.nan)
    out = out.cpu().numpy()
    for r in range(world):
        for j, c in enumerate(assign(len(candidates), world, r)):
            res[c] = out[r, j]
    return res



class PopulationEvaluator:
    """Fitness of a whole population of lambda candidates against ONE frozen supernet (the reference walks the population
    strictly sequentially, sasrec/evolution.py:172-206, re-sampling the validation negatives on the CPU for every candidate).

      * the validation batches (sequences + 1 + C sampled candidates per user) are assembled ONCE on the device
        (adt_b200.sampler.DeviceSampler.eval_batch) and shared by every candidate -- every candidate is ranked on the same negatives;
      * per candidate the host only flips `set_choice` and enqueues the encoder blocks + ONE gather-dot / rank / metric launch per
        batch (adt_candidate_scores): AUC / NDCG@10 / HR@10 accumulate on the device, nothing is read back per batch;
      * `in_flight` candidates are enqueued on separate CUDA streams: their evaluations are independent given the frozen weights, and
        one candidate's latency-bound kernels leave most of a B200 idle;
      * with a process group the candidates are dealt round-robin to the ranks (candidate c -> rank c mod G) and the fitness triples
        are all-gathered (evaluate_population)."""

    def __init__(self, model, val_batches, rec_choice, ind_choice, process_group=None, in_flight=4):
        self.model, self.batches = model, list(val_batches)
        self.rec_choice, self.ind_choice = rec_choice, ind_choice
        self.pg = process_group
        dev = model.item_emb.weight.device
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, int(in_flight)))]
        self.evaluated = 0

    @torch.no_grad()
    def _enqueue(self, cand, stream):
        dev = self.model.item_emb.weight.device
        acc = torch.zeros(7, dtype=torch.float64, device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            set_choice_from_candidate(self.model, cand, self.rec_choice, self.ind_choice)
            for seq, item_idx in self.batches:
                sampled_rank(self.model, seq, item_idx, metric_acc=acc)
        return acc

    @torch.no_grad()
    def fitness_many(self, cands):
        """[(AUC, NDCG@10, HR@10)] for this process's candidates, `in_flight` of them enqueued at a time"""
        self.model.eval()
        out, pending = [], []
        for i, cand in enumerate(cands):
            pending.append(self._enqueue(cand, self.streams[i % len(self.streams)]))
            if len(pending) == len(self.streams) or i == len(cands) - 1:
                for s in self.streams:
                    torch.cuda.current_stream().wait_stream(s)
                for acc in pending:
                    (ndcg, hr), auc, _ = sampled_metrics_from_acc(acc)
                    out.append((auc, ndcg[10], hr[10]))
                pending = []
        self.evaluated += len(cands)
        return out

    def fitness(self, cand):
        return self.fitness_many([cand])[0]

    def evaluate(self, candidates):
        """-> array [n_candidates, 3] identical on every rank"""
        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
        world = torch.distributed.get_world_size(self.pg) if dist_on else 1
        rank = torch.distributed.get_rank(self.pg) if dist_on else 0
        mine = assign(len(candidates), world, rank)
        return evaluate_population(candidates, None, self.pg, local=self.fitness_many([candidates[c] for c in mine]))
