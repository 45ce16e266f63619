"""`main_missing.py` of the reference on the rd_b200 hot path — same config.yaml, same flow, same outputs.

    python -m rd_b200.main_missing [--config config.yaml] [--synthetic SUBJECTS] [--set key=value ...]

Mirrors the reference entry script (`src/main_missing.py`, a module-level script; cited line by line below):
seeds (:17-22), config load + derived keys + checkpoint directory / saved-yaml override (:25-56), data loaders (:61-68),
model construction (:71-95), `fix_pretrain` freezing (:104-116), Adam(amsgrad, wd 1e-5) + ReduceLROnPlateau (:118-119),
checkpoint restore by key with the shape-filtered model load (:125-135, src/util.py:870-903), `train()` (:141-335) with the
per-epoch loss averages, `stat.csv` rows, validation, scheduler step on the monitored metric and `epochNNN.pth.tar` /
`model_best.pth.tar` checkpoints, and `evaluate()` (:337-609) with the reconstruction / segmentation metrics.

What differs is where the work runs: every network, loss, the clip and the optimizer are rd_b200 kernels (one CUDA graph per
iteration), the metrics are computed on the device, batches come from a device-resident volume store (rd_b200.data.VolumeStore /
SlabLoader; `--synthetic N` builds a random one, `data_factory` lets a caller plug real volumes).  Out of scope:
HDF5 / NIfTI readers, result dumps to H5, the nearest-neighbour latent search — the hooks are `data_factory` and `save_res`.
"""
import argparse
import copy
import os
import shutil
import sys
import time

import numpy as np
import torch
import yaml

from . import config as rd_config
from . import data as rd_data
from .trainer import LOSS_KEYS, Trainer, apply_fix_pretrain, build_model

STAT_KEYS = ["recon_y", "recon_y_fused", "recon_x", "recon_x_mix", "kl", "latent_z", "sim_s", "sim_z", "adv_s", "adv_s_d", "all"]


# ------------------------------------------------------------------------------------------------ util.py equivalents
def save_config_yaml(ckpt_path, config):
    """src/util.py:915-925: only plain values are written."""
    plain = {k: v for k, v in config.items() if isinstance(v, (int, float, str, list, dict))}
    with open(os.path.join(ckpt_path, "config.yaml"), "w") as f:
        yaml.dump(copy.deepcopy(plain), f)


def save_config_file(config):
    """src/util.py:846-851."""
    with open(os.path.join(config["ckpt_path"], "config.txt"), "w") as f:
        for k, v in config.items():
            f.write(k + ": " + str(v) + "\n")


def save_result_stat(stat, config, info="Default"):
    """src/util.py:854-866: one CSV row per call, columns = 'info' + sorted stat keys (written without pandas)."""
    path = os.path.join(config["ckpt_path"], "stat.csv")
    cols = ["info"] + sorted(k for k in stat.keys() if k != "info")
    new = not os.path.exists(path)
    with open(path, "a") as f:
        if new:
            f.write("," + ",".join(cols) + "\n")
        row = dict(stat)
        row["info"] = info
        f.write("0," + ",".join(str(row[c]) for c in cols) + "\n")


def save_checkpoint(state, is_best, checkpoint_dir):
    """src/util.py:148-153."""
    filename = os.path.join(checkpoint_dir, "epoch" + str(state["epoch"]).zfill(3) + ".pth.tar")
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, os.path.join(checkpoint_dir, "model_best.pth.tar"))


def load_checkpoint_model(model, pretrained_dict):
    """src/util.py:896-903: keys that exist with the same shape are taken, everything else keeps its value."""
    model_dict = model.state_dict()
    model_dict.update({k: v for k, v in pretrained_dict.items() if k in model_dict and v.shape == model_dict[k].shape})
    model.load_state_dict(model_dict)
    return model


class _LRShim:
    """ReduceLROnPlateau needs an optimizer object; the real optimizer is the fused kernel of the Trainer, so the scheduler steers a
    one-parameter torch optimizer whose learning rate is copied into the Trainer's hyper-parameter buffer after every step."""

    def __init__(self, trainer: Trainer, lr: float):
        self.trainer = trainer
        self.opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=lr, weight_decay=1e-5, amsgrad=True)
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.opt, mode="min", factor=0.1, patience=5, min_lr=1e-5)

    def step(self, metric):
        self.scheduler.step(metric)
        self.trainer.set_lr(self.opt.param_groups[0]["lr"])


class Run:
    """State of one `main_missing.py` execution (the reference keeps all of this in module globals)."""

    def __init__(self, config: dict, data_factory=None, device=None, log=print):
        self.log = log
        # ---- seeds (:17-22)
        seed = 10
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(seed)
        np.random.seed(seed)
        # ---- config (:25-56)
        config = rd_config.derive(config)
        dev = device if device is not None else torch.device("cuda:" + str(config["gpu"]))
        config["device"] = dev
        if config["ckpt_timelabel"] and (config["phase"] == "test" or config["continue_train"] is True):
            time_label = config["ckpt_timelabel"]
        else:
            lt = time.localtime(time.time())
            time_label = "%d_%d_%d_%d_%d" % (lt.tm_year, lt.tm_mon, lt.tm_mday, lt.tm_hour, lt.tm_min)
        config["ckpt_path"] = os.path.join(config.get("ckpt_root", "../ckpt/"), config["dataset_name"], config["model_name"], time_label)
        if not os.path.exists(config["ckpt_path"]):
            os.makedirs(config["ckpt_path"])
            save_config_yaml(config["ckpt_path"], config)
        elif config["load_yaml"]:
            ypath = os.path.join(config["ckpt_path"], "config.yaml")
            if os.path.exists(ypath):
                with open(ypath) as f:
                    loaded = yaml.safe_load(f)
                for k, v in loaded.items():
                    if k in ("phase", "continue_train"):
                        continue
                    if k in config:
                        config[k] = v
                config = rd_config.derive(config)
                config["device"] = dev
            else:
                save_config_yaml(config["ckpt_path"], config)
        if config["model_name"] != "MultimodalModel":
            raise ValueError("not supporting other models yet!")
        self.config = config
        # ---- data (:61-68)
        factory = data_factory or synthetic_data_factory
        self.train_loader, self.val_loader, self.test_loader = factory(config, dev)
        # ---- model + freezing + optimizer (:71-119)
        self.model = build_model(config, dev)
        if apply_fix_pretrain(self.model, config):
            self.log("----------------Fixed stage 1 parts!------------------")
        self.trainer = Trainer(self.model, config, config["batch_size"], use_graph=bool(config.get("cuda_graph", True)) and dev.type == "cuda")
        self.sched = _LRShim(self.trainer, config["lr"])
        # ---- restore (:125-137)
        self.start_epoch = -1
        if config["continue_train"] or config["phase"] == "test":
            self.start_epoch = self.load_checkpoint(config["ckpt_name"])
        if config["phase"] == "train":
            save_config_file(config)

    # ------------------------------------------------------------------ checkpoint by key (src/util.py:870-893)
    def load_checkpoint(self, ckpt_name):
        filename = os.path.join(self.config["ckpt_path"], ckpt_name)
        if not os.path.isfile(filename):
            raise ValueError("No correct checkpoint")
        ck = torch.load(filename, map_location="cpu", weights_only=False)
        for key in ("optimizer", "scheduler", "model"):
            try:
                if key == "model":
                    load_checkpoint_model(self.model, ck[key])
                elif key == "optimizer":
                    self.trainer.load_optimizer_state_dict(ck[key])
                    self.sched.opt.param_groups[0]["lr"] = self.trainer.get_lr()
                else:
                    self.sched.scheduler.load_state_dict(ck[key])
                self.log("loading " + key + " success!")
            except Exception:
                self.log("loading " + key + " failed!")
        self.log("loaded checkpoint from '%s' (epoch: %s, monitor metric: %s)" % (filename, ck["epoch"], ck.get("monitor_metric")))
        return ck["epoch"]

    # ------------------------------------------------------------------ train() (:141-335)
    def train(self):
        cfg, tr = self.config, self.trainer
        global_iter = 0
        monitor_metric_best = 100
        stat = None
        for epoch in range(self.start_epoch + 1, cfg["epochs"]):
            self.model.train()
            tr.start_epoch()
            acc = torch.zeros(len(LOSS_KEYS), dtype=torch.float64, device=tr.dev)
            global_iter0 = global_iter
            for it, sample in enumerate(self.train_loader, 0):
                global_iter += 1
                loss_vec = tr.train_iteration(sample, with_y=(it == 0))       # iter 0: y is decoded once even without a y-loss (:182-185)
                acc += loss_vec.double()                                      # device-side sum: no .item() per term
                if global_iter % 10 == 0:
                    v = dict(zip(LOSS_KEYS, loss_vec.tolist()))
                    self.log("Epoch[%3d], iter[%3d]: loss=[%.4f], recon x=[%.4f], recon x_mix=[%.4f], recon y=[%.4f], recon y_fused=[%.4f], "
                             "kl=[%.4f], latent z=[%.4f], sim s=[%.4f], sim z=[%.4f], adv s=[%.4f], adv s d=[%.4f]"
                             % (epoch, it, v["all"], v["recon_x"], v["recon_x_mix"], v["recon_y"], v["recon_y_fused"], v["kl"],
                                v["latent_z"], v["sim_s"], v["sim_z"], 0.0, 0.0))
                if cfg.get("max_iters") and it + 1 >= cfg["max_iters"]:
                    break
            num_iter = max(global_iter - global_iter0, 1)
            loss_all = {k: 0.0 for k in STAT_KEYS}
            loss_all.update(dict(zip(LOSS_KEYS, (acc / num_iter).tolist())))
            save_result_stat(dict(loss_all), cfg, info="epoch[%2d]" % epoch)
            self.log(loss_all)
            stat = self.evaluate(phase="val", set="val", save_res=False)
            if cfg["lambda_recon_y"] == 0 or cfg["lambda_recon_y_fused"] == 0:
                monitor_metric = stat["recon_x_mix"]
            else:
                monitor_metric = stat["recon_y_fused"]
            self.sched.step(monitor_metric)
            save_result_stat(dict(stat), cfg, info="val")
            self.log(stat)
            is_best = monitor_metric <= monitor_metric_best
            if is_best:
                monitor_metric_best = monitor_metric
            state = {"epoch": epoch, "monitor_metric": monitor_metric, "stat": stat, "optimizer": tr.optimizer_state_dict(),
                     "scheduler": self.sched.scheduler.state_dict(), "model": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()}}
            save_checkpoint(state, is_best, cfg["ckpt_path"])
        return stat

    # ------------------------------------------------------------------ evaluate() (:337-609)
    def evaluate(self, phase="val", set="val", save_res=True, info=""):
        cfg, tr = self.config, self.trainer
        self.model.eval()
        if phase == "val":
            loader = self.val_loader
        elif set == "train":
            loader = self.train_loader
        elif set == "val":
            loader = self.val_loader
        elif set == "test":
            loader = self.test_loader
        else:
            raise ValueError("Undefined loader")
        acc = torch.zeros(len(LOSS_KEYS), dtype=torch.float64, device=tr.dev)
        metric_rows = {}
        kept = {"inputs": [], "targets": [], "mask": [], "y_fake_fused": [], "xi_fake": [], "xi_fake_mix": [], "s_list": [], "z_list": [],
                "subj_id": [], "slice_idx": []}
        res_path = os.path.join(cfg["ckpt_path"], "result_" + set)
        os.makedirs(res_path, exist_ok=True)
        n_iter = 0
        y_on = cfg["lambda_recon_y"] > 0 or cfg["lambda_recon_y_fused"] > 0
        for it, sample in enumerate(loader, 0):
            loss_vec, metrics, T = tr.eval_iteration(sample, with_y=(it == 0 or y_on))
            acc += loss_vec.double()
            for k, v in metrics.items():
                metric_rows.setdefault(k, []).append(v)
            if phase == "test" and save_res:
                B, M = int(sample["inputs"].shape[0]), tr.M
                kept["inputs"].append(sample["inputs"].detach().float().cpu())
                kept["targets"].append(sample["targets"].detach().float().cpu())
                kept["mask"].append(sample["mask"].detach().float().cpu())
                if T["y_fake_fused"] is not None:
                    kept["y_fake_fused"].append(T["y_fake_fused"].permute(0, 3, 1, 2).float().cpu())
                kept["xi_fake"].append(T["x_fake"].reshape(M, B, *T["x_fake"].shape[1:]).permute(1, 0, 4, 2, 3).float().cpu())
                kept["xi_fake_mix"].append(T["x_fake_mix"].reshape(M * (M - 1), B, *T["x_fake_mix"].shape[1:]).permute(1, 0, 4, 2, 3).float().cpu())
                kept["s_list"].append(T["S"].reshape(M, B, *T["S"].shape[1:]).permute(1, 0, 4, 2, 3).float().cpu())
                kept["z_list"].append(T["z"].reshape(M, B, -1).permute(1, 0, 2).float().cpu())
                kept["subj_id"] += list(sample.get("subj_id", []))
                kept["slice_idx"].append(torch.as_tensor(sample.get("slice_idx", torch.zeros(B))).cpu())
            n_iter = it + 1
            if it > 500 or (cfg.get("max_eval_iters") and n_iter >= cfg["max_eval_iters"]):
                break
        stat = {k: 0.0 for k in STAT_KEYS}
        stat.update(dict(zip(LOSS_KEYS, (acc / max(n_iter, 1)).tolist())))
        for k, rows in metric_rows.items():       # ONE device -> host read per metric for the whole evaluation
            stat[k] = float(np.array(torch.cat(rows).double().cpu().tolist()).mean()) if rows else float("nan")
        if phase == "test" and save_res and kept["inputs"]:
            # the reference writes results_all<info>.h5 (h5py is absent here); same keys, torch.save container
            out = {k: (torch.cat(v, 0) if v and torch.is_tensor(v[0]) else v) for k, v in kept.items() if v}
            torch.save(out, os.path.join(res_path, "results_all" + info + ".pt"))
        return stat


def synthetic_data_factory(config, device):
    """Train / val / test SlabLoaders over a random device-resident VolumeStore with the reference's batch layout (the reference's
    HDF5 files are not available; a caller with real volumes passes its own factory returning three iterables of batch dicts)."""
    n_subj = int(config.get("synthetic_subjects", 4))
    D = int(config.get("synthetic_depth", 24))
    store = rd_data.VolumeStore.synthetic(n_subj, config["contrast_list"], config["dataset_name"], D=D, H=config["input_height"],
                                          W=config["input_width"], seed=10, device=device,
                                          missing_prob=float(config.get("synthetic_missing", 0.0)))
    bs = config["block_size"]
    slices = list(range(bs, D - bs))
    split = max(1, n_subj // 4)
    def pairs(subjects):
        return [store.subj_ids[s] for s in subjects for _ in slices], [i for _ in subjects for i in slices]
    tr_s, tr_i = pairs(range(0, max(1, n_subj - 2 * split)))
    va_s, va_i = pairs(range(max(1, n_subj - 2 * split), max(2, n_subj - split)) if n_subj > 2 else range(n_subj))
    te_s, te_i = pairs(range(max(2, n_subj - split), n_subj) if n_subj > 2 else range(n_subj))
    mk = lambda s, i, sh, dr: rd_data.SlabLoader(store, s, i, config["batch_size"], shuffle=sh, dropoff=dr, block_size=bs)
    return mk(tr_s, tr_i, config["shuffle"], config["dropoff"]), mk(va_s, va_i, False, config["dropoff"]), mk(te_s, te_i, False, False)


def main(argv=None):
    ap = argparse.ArgumentParser(description="representation-disentanglement main_missing.py on the rd_b200 kernels")
    ap.add_argument("--config", default="config.yaml")
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic subjects (device-resident random volumes)")
    ap.add_argument("--set", action="append", default=[], help="key=value config override (YAML value syntax)")
    a = ap.parse_args(argv)
    over = {}
    for kv in a.set:
        k, v = kv.split("=", 1)
        over[k] = yaml.safe_load(v)
    if a.synthetic:
        over["synthetic_subjects"] = a.synthetic
    cfg = rd_config.load_config(a.config, **over) if os.path.exists(a.config) else rd_config.default_config(**over)
    run = Run(cfg)
    if cfg["phase"] == "train":
        stat = run.train()
    else:
        stat = run.evaluate(phase="test", set="test", save_res=True)
    print(stat)
    return stat


if __name__ == "__main__":
    main()
