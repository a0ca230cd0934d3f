"""Result / log / checkpoint formats of the reference's drivers, so that the reference's tooling
(candidates_to_lambdas.py, the templates, `--state_dict_path`) keeps working on what this package writes.

  log_line()            sasrec/main.py:202            `str(t_valid) + ' ' + str(t_test)`  with t = (NDCG dict, HT dict)
  checkpoint_name()     sasrec/main.py:208, :216-217  SASRec.epoch=..lr=..layer=..head=..hidden=..maxlen=...pth (main.py:105-107 parses `epoch=`)
  res_jsonl_name()      sasrec/evolution.py:355       ./res/res_<dataset>_lr_.._population_.._mutation_...jsonl
  write_res_jsonl()     sasrec/evolution.py:355-363   one JSON object per kept candidate: the vis_dict entry + 'cand', 'rec', 'ind' (as str)
  read_res_jsonl()      the README workflow: best candidate -> lambdas (candidates_to_lambdas.py)
State dicts themselves are plain torch.save(model.state_dict()) with the reference's parameter names (tests/test_host_cpu.py).
"""
import json
import os

from .lambdas import candidate_to_lambdas


def log_line(t_valid, t_test):
    """t_* = (NDCG: {k: float}, HT: {k: float}) as evaluate_loader returns them"""
    return str(t_valid) + " " + str(t_test) + "\n"


def parse_log_line(line):
    """inverse of log_line (the dict reprs are Python literals)"""
    import ast
    line = line.strip()
    depth, cut = 0, None
    for i, ch in enumerate(line):
        depth += ch == "("
        depth -= ch == ")"
        if depth == 0 and ch == ")":
            cut = i + 1
            break
    return ast.literal_eval(line[:cut]), ast.literal_eval(line[cut:].strip())


def checkpoint_name(epoch, lr, num_layers, num_heads, hidden_units, maxlen):
    return f"SASRec.epoch={epoch}.lr={lr}.layer={num_layers}.head={num_heads}.hidden={hidden_units}.maxlen={maxlen}.pth"


def epoch_from_checkpoint(path):
    """main.py:105-107: `tail = path[path.find('epoch=') + 6:]; int(tail[:tail.find('.')]) + 1`"""
    tail = path[path.find("epoch=") + 6:]
    return int(tail[:tail.find(".")]) + 1


def res_jsonl_name(args):
    return (f"./res/res_{args.dataset}_lr_{args.lr}_reg_{args.weight_decay}_warm_{args.warmup_epochs}_search_{args.search_epochs}"
            f"_layers_{args.num_layers}_select_{args.select_num}_population_{args.population_num}_cross_{args.crossover_num}"
            f"_mutation_{args.mutation_num}.jsonl")


def write_res_jsonl(path, top_candidates, vis_dict, rec_choice, ind_choice):
    """evolution.py:355-363.  top_candidates: the kept candidate vectors (lists of floats in [0,1]); vis_dict[str(cand)] holds
    'visited', 'auc', 'V_NDCG', 'V_HR', 'V_AUC' (evolution.py:172-190)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        for cand in top_candidates:
            info = dict(vis_dict[str(cand)])
            rec, ind = candidate_to_lambdas(list(cand), rec_choice, ind_choice)
            info["cand"] = str(cand)
            info["rec"] = str(rec)
            info["ind"] = str(ind)
            f.write(json.dumps(info) + "\n")


def read_res_jsonl(path):
    import ast
    out = []
    with open(path) as f:
        for line in f:
            if line.strip():
                d = json.loads(line)
                for k in ("cand", "rec", "ind"):
                    if k in d:
                        d[k] = ast.literal_eval(d[k])
                out.append(d)
    return out
