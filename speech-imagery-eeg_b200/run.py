"""Entry point with the reference's flag surface (run.py:14-144): loops seeds, builds Experiment(args),
trains unless a checkpoint exists, tests, pickles the result.  Launch one process per GPU with torchrun
for data-parallel training.  `--data synthetic` (default when no dataset directory exists) draws seeded
synthetic series of the named dataset's shape."""
import argparse
import os
import pickle
import random

import numpy as np
import torch

FLAGS = [
    # name, type/action, default
    ("--data", str, "synthetic"), ("--data_root", str, "./data/UEA_multivariate"), ("--json_path", str, ""),
    ("--target_channels", int, 122), ("--target_timepoints", int, 1651), ("--max_files", int, 1000),
    ("--max_subjects", int, 5), ("--subject_id", str, "sub-01"), ("--task_type", str, "imagine"),
    ("--model", str, "InterpGN"), ("--dnn_type", str, "Transformer"), ("--dataset", str, "BasicMotions"),
    ("--lambda_reg", float, 0.1), ("--lambda_div", float, 0.1), ("--epsilon", float, 1.), ("--num_shapelet", int, 10),
    ("--gating_value", float, None), ("--pos_weight", "store_true", False), ("--sbm_cls", str, "linear"),
    ("--distance_func", str, "euclidean"), ("--beta_schedule", str, "constant"), ("--memory_efficient", "store_true", False),
    ("--lr", float, 5e-3), ("--lr_decay", "store_true", False), ("--gradient_accumulation_steps", int, 1),
    ("--gradient_clip", float, 0), ("--batch_size", int, 64), ("--log_interval", int, 20), ("--min_epochs", int, 0),
    ("--train_epochs", int, 500), ("--num_workers", int, 0), ("--patience", int, 50), ("--multi_gpu", "store_true", False),
    ("--test_only", "store_true", False), ("--seed", int, -1), ("--amp", "store_false", True),
    ("--task_name", str, "classification"), ("--model_id", str, "test"), ("--embed", str, "timeF"), ("--freq", str, "h"),
    ("--top_k", int, 5), ("--num_kernels", int, 6), ("--enc_in", int, 7), ("--dec_in", int, 7), ("--c_out", int, 7),
    ("--d_model", int, 512), ("--n_heads", int, 8), ("--e_layers", int, 2), ("--d_layers", int, 1), ("--d_ff", int, 2048),
    ("--moving_avg", int, 25), ("--factor", int, 1), ("--distil", "store_false", True), ("--dropout", float, 0),
    ("--activation", str, "gelu"), ("--output_attention", "store_true", False), ("--label_len", int, 48),
    ("--pred_len", int, 96), ("--seasonal_patterns", str, "Monthly"), ("--inverse", "store_true", False),
    # new in this framework
    ("--shapelet_precision", str, "fp32"), ("--syn_shape", str, None), ("--syn_train", int, 512),
    ("--syn_val", int, 128), ("--syn_test", int, 128), ("--syn_subjects", int, 1),
    ("--loso", "store_true", False),      # leave-one-subject-out: one fold per subject, folds spread over the ranks
    ("--cuda_graph", "store_true", False),        # replay each full-batch training step from one captured CUDA graph
    ("--allow_synthetic", "store_true", False),   # --data UEA|EEG|EEG3 without the archive on disk: synthetic series of its shape
]


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    for name, kind, default in FLAGS:
        if kind in ("store_true", "store_false"):
            parser.add_argument(name, action=kind, default=default)
        else:
            parser.add_argument(name, type=kind, default=default)
    parser.add_argument("--subject_ids", type=str, nargs='+', default=["sub-01,sub-02,sub-03"])
    args = parser.parse_args(argv)
    if args.data not in ("synthetic", "EEG", "EEG3", "UEA"):
        parser.error("--data must be one of synthetic, EEG, EEG3, UEA")
    args.root_path = args.data_root if args.data in ('EEG', 'EEG3') else f"{args.data_root}/{args.dataset}"
    # where the samples really come from; it is part of the checkpoint directory name and of test_results.pkl.
    # A real archive whose directory is missing raises in data_provider unless --allow_synthetic is given.
    args.data_source = "synthetic" if (args.data == "synthetic" or not os.path.isdir(args.root_path)) else args.data
    args.is_training = True
    return args


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def run_loso(args):
    """Leave-one-subject-out (exp/loso.py): fold f trains on rank f % world, results are gathered at the end."""
    from data_provider.data_factory import collate_fn, loso_dataset
    from exp import loso
    from exp.experiment_classification import Experiment
    from exp.parallel import init_distributed
    import datetime
    import json
    # ranks own different numbers of folds and meet only at the final gather: a rank that finishes early (or owns no
    # fold at all) must be allowed to wait for the slowest one, far beyond NCCL's default 10-minute watchdog
    rank, _, world = init_distributed(timeout=datetime.timedelta(days=7))
    seed = 0 if args.seed == -1 else args.seed
    set_seed(seed)
    args.seed = seed
    ds = loso_dataset(args)
    folds = loso.loso_folds(ds.subject, seed=seed)
    mine, local = loso.folds_of_rank(len(folds), world, rank), {}
    for f in mine:
        subject = folds[f][0]
        tr, va, te = loso.fold_loaders(ds, folds[f], args.batch_size, collate_fn, args.num_workers,
                                       pin_memory=torch.cuda.is_available())
        args.dataset_tag = f"loso-s{subject}"
        exp = Experiment(args, loaders=((ds, tr), (ds, va), (ds, te)), independent=True)
        exp.checkpoint_dir = os.path.join(exp.checkpoint_dir, f"loso_subject_{subject}")
        print(f"[rank {rank}] fold {f + 1}/{len(folds)}: held-out subject {subject}, "
              f"{len(folds[f][1])} train / {len(folds[f][2])} val / {len(folds[f][3])} test samples", flush=True)
        exp.train()
        loss, res, _ = exp.test(save_csv=False)
        local[f] = (subject, loss, None if res is None else res.accuracy)
        os.makedirs(exp.checkpoint_dir, exist_ok=True)      # a fold's result survives a later failure of another rank
        with open(os.path.join(exp.checkpoint_dir, "fold_result.json"), "w") as fh:
            json.dump({"fold": f, "subject": int(subject), "test_loss": float(loss), "accuracy": local[f][2],
                       "rank": rank}, fh)
        torch.cuda.empty_cache()
    merged = loso.gather_results(local, world)
    mean_acc, rows = loso.summarize(merged)
    if rank == 0:
        for subject, loss, acc in rows:
            print(f"LOSO subject {subject}: test loss {loss:.4f}  accuracy {acc if acc is None else round(100 * acc, 2)}%")
        print(f"LOSO mean accuracy over {len(rows)} subjects: {100 * mean_acc:.2f}%")
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return merged


def main(argv=None):
    args = get_args(argv)
    if args.task_name != "classification":
        raise SystemExit("only --task_name classification is built on the B200 hot path")
    if args.loso:
        return run_loso(args)
    from exp.experiment_classification import Experiment
    seeds = [0, 42, 1234, 8237, 2023] if args.seed == -1 else [args.seed]
    results = []
    for i, seed in enumerate(seeds):
        set_seed(seed)
        args.seed = seed
        exp = Experiment(args)
        exp.log(f"===== run {i + 1}/{len(seeds)}  seed {seed}  model {args.model}/{args.dnn_type}  "
                f"shape (C={args.enc_in}, T={args.seq_len}, classes={args.num_class})  world {exp.world} =====")
        ckpt = os.path.join(exp.checkpoint_dir, "checkpoint.pth")
        if not args.test_only:
            if os.path.exists(ckpt):
                exp.log(f"checkpoint exists, skipping training: {ckpt}")
            else:
                exp.train()
                torch.cuda.empty_cache()
        if os.path.exists(ckpt):
            exp.model.load_state_dict(torch.load(ckpt, map_location=exp.device))
        elif args.test_only:
            exp.log(f"no checkpoint at {ckpt}; testing the randomly initialised model")
        loss, res, df = exp.test(save_csv=True, result_dir=f"./result/{args.model}")
        results.append((loss, None if res is None else res.accuracy))
        if exp.is_main and res is not None:
            with open(os.path.join(exp.checkpoint_dir, "test_results.pkl"), "wb") as f:
                pickle.dump({"test_loss": loss, "test_metrics": res, "test_df": df, "args": vars(args)}, f)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return results


if __name__ == "__main__":
    main()
