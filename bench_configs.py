#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one B200, device-resident
inputs, CUDA-event timing.  `bench.py` is the contract benchmark (config 3) and attaches short legs of these
configurations to its JSON line (`configs`); this script runs the same legs (`bench.config_leg`) with more
steps and any subset, one JSON line per configuration.

    python bench_configs.py [--configs cfg1,cfg2,cfg3_f64,cfg3_full,cfg4,cfg5] > profiles/configs_rNN.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import CFG2_OPS, config_leg  # noqa: E402,F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg1_f32,cfg2,cfg2_f64,cfg3_f64,cfg3_full,cfg4,cfg4_f64,cfg5,cfg5_f64")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    for name in args.configs.split(","):
        res = config_leg(name, torch, dev, args.steps, args.warmup)
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
