"""Appends the reference's optimizer layout to tests/golden/state_dict_reference.json: the parameter groups
trainval_net.py:280-294 builds for the reference's own class-agnostic module under cfgs/res101.yml (group order = parameter
names, lr and weight decay per group).  This container only.

    python tests/golden/make_golden_optimizer.py
"""
import json
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ns = runpy.run_path(os.path.join(HERE, "make_golden_state_dict.py"))          # builds the reference module (and rewrites the json)
cfg, resnet = ns["cfg"], ns["resnet"]
cfg.TRAIN.WEIGHT_DECAY, cfg.TRAIN.DOUBLE_BIAS, cfg.TRAIN.LEARNING_RATE = 0.0001, False, 0.001   # cfgs/res101.yml:11-13
net = resnet(tuple(range(31)), 101, pretrained=False, class_agnostic=True)
net.create_architecture()
lr = 0.001                                                                      # trainval_net.py:88-90 (--lr default)
groups = []
for key, value in dict(net.named_parameters()).items():                        # trainval_net.py:280-287, verbatim semantics
    if value.requires_grad:
        if 'bias' in key:
            groups.append([key, lr * (cfg.TRAIN.DOUBLE_BIAS + 1), cfg.TRAIN.BIAS_DECAY and cfg.TRAIN.WEIGHT_DECAY or 0])
        else:
            groups.append([key, lr, cfg.TRAIN.WEIGHT_DECAY])
path = os.path.join(HERE, "state_dict_reference.json")
blob = json.load(open(path))
blob["optimizer_groups_res101_yml"] = groups
json.dump(blob, open(path, "w"))
print(len(groups), "parameter groups")
