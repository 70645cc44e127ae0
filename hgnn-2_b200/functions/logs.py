"""Minimal stand-in for the reference's functions/logs.py (out of scope: SURVEY.md section 2 #14).

scripts/train_mnb.py:21 and test_mnb.py:21 import ``logs`` from ``functions``, and the drivers call a
``Logger``; the reference's version needs matplotlib (absent offline) and draws plots.  This one
keeps text logging and whole-module save/load (logs.py:99-123) and accepts every ``add_*`` /
``plot_*`` call as a recorded no-op so the unmodified drivers run.
"""
import os

import torch


class Logger(object):
    def __init__(self, log_dir):
        self.log_dir = log_dir
        self.records = []
        self.time_epoch = []
        os.makedirs(os.path.join(log_dir, "parameters"), exist_ok=True)

    def write_settings(self, args):
        with open(os.path.join(self.log_dir, "experiment.txt"), "w") as f:
            for k, v in sorted(vars(args).items()):
                f.write("%s : %s\n" % (k, v))

    def save_model(self, model):
        torch.save(model, os.path.join(self.log_dir, "parameters", "gnn.pt"))

    def load_model(self, dpath):
        return torch.load(os.path.join(dpath, "parameters", "gnn.pt"), weights_only=False)

    def add_epoch_info(self, epoch, loss, error, dur):
        self.time_epoch.append(dur)
        self.records.append(("epoch", epoch, loss, error, dur))

    def __getattr__(self, name):
        if name.startswith(("add_", "plot_", "write_")):
            return lambda *a, **k: self.records.append((name,) + a)
        raise AttributeError(name)
