"""Per-dataset disentanglement weights and the search-space mapping.

get_lambdas  mirrors /root/reference/sasrec/utils.py:850-862 (same tables, same return order: reconstruction
             lambdas1 per block, independence lambdas2 per block).
get_weight   mirrors /root/reference/sasrec/evolution.py:124-137 and /root/reference/candidates_to_lambdas.py:3-9
             (piece-wise linear interpolation of a candidate in [0,1] on the choice grid).
"""

_TABLE = {
    "ml-1m": ([0.104292, 0.065892], [0.100833, 0.000607]),
    "beauty": ([0.0124, 0.122], [0.0001, 0.0]),
    "Beauty": ([0.0124, 0.122], [0.0001, 0.0]),
    "steam": ([0.0001, 0.0005], [0.00134, 0.00028]),
    "ml-20m": ([0.005, 0.1], [0.00186667, 0.075]),
}


def get_lambdas(dataset, tp=-1):
    if dataset not in _TABLE:
        return None  # the reference falls through and returns None for unknown datasets
    l1, l2 = _TABLE[dataset]
    return list(l1), list(l2)


def get_weight(choices, prob):
    split_value = 1 / (len(choices) - 1)
    idx = 0
    while prob > split_value:
        idx += 1
        prob -= split_value
    relate_distance = prob / split_value
    return choices[idx] * (1 - relate_distance) + choices[idx + 1] * relate_distance


def candidate_to_lambdas(cand, rec_choice, ind_choice):
    """sasrec/evolution.py:139-153: [rec0, ind0, rec1, ind1, ...] in [0,1] -> (rec_weights, ind_weights)."""
    n = len(cand) // 2
    rec = [get_weight(rec_choice, cand[2 * i]) for i in range(n)]
    ind = [get_weight(ind_choice, cand[2 * i + 1]) for i in range(n)]
    return rec, ind
