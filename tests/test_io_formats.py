"""log.txt / res/*.jsonl / checkpoint-name formats of the reference's drivers."""
import os
import subprocess
import sys
import types

import numpy as np


def test_log_line_round_trip_and_reference_shape():
    from adt_b200.io_formats import log_line, parse_log_line
    t_valid = ({5: 0.1234, 10: 0.2345}, {5: 0.3, 10: 0.45})
    t_test = ({5: 0.11, 10: 0.22}, {5: 0.33, 10: 0.44})
    line = log_line(t_valid, t_test)
    assert line == str(t_valid) + " " + str(t_test) + "\n"          # sasrec/main.py:202 verbatim
    assert parse_log_line(line) == (t_valid, t_test)


def test_checkpoint_name_is_parsed_like_main_py():
    from adt_b200.io_formats import checkpoint_name, epoch_from_checkpoint
    name = checkpoint_name(17, 0.001, 2, 2, 64, 50)
    assert name == "SASRec.epoch=17.lr=0.001.layer=2.head=2.hidden=64.maxlen=50.pth"
    assert epoch_from_checkpoint("/x/y/" + name) == 18                # main.py:105-107


def test_res_jsonl_feeds_candidates_to_lambdas(tmp_path):
    """the file evolution.py:355-363 writes: str-encoded cand / rec / ind next to the fitness record; the rec / ind columns must be what
    the reference's own candidates_to_lambdas.py computes for that candidate."""
    from adt_b200.io_formats import res_jsonl_name, write_res_jsonl, read_res_jsonl
    rec_choice = [0, 0.0001, 0.0005, 0.001, 0.005, 0.01]
    ind_choice = [0, 0.0001, 0.0005, 0.001, 0.005, 0.01]
    cand = [0.7053411308078107, 0.9542592593410837, 0.9296478828883573, 0.28425047269448145, 0.1600125621449342, 0.47495464861462977]
    vis = {str(cand): {"visited": True, "auc": 0.91, "V_NDCG": 0.3, "V_HR": 0.5, "V_AUC": 0.91}}
    args = types.SimpleNamespace(dataset="beauty", lr=0.001, weight_decay=0.0, warmup_epochs=5, search_epochs=3, num_layers=3, select_num=5,
                                 population_num=10, crossover_num=3, mutation_num=3)
    name = res_jsonl_name(args)
    assert name == "./res/res_beauty_lr_0.001_reg_0.0_warm_5_search_3_layers_3_select_5_population_10_cross_3_mutation_3.jsonl"
    path = os.path.join(tmp_path, os.path.basename(name))
    write_res_jsonl(path, [cand], vis, rec_choice, ind_choice)
    rows = read_res_jsonl(path)
    assert rows[0]["cand"] == cand and rows[0]["auc"] == 0.91 and rows[0]["visited"] is True
    golden = open(os.path.join(os.path.dirname(__file__), "golden", "candidates_to_lambdas.txt")).read().strip()
    assert f"{rows[0]['rec']} {rows[0]['ind']}" == golden              # the reference script's printed output for this candidate
