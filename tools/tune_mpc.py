"""Sweeps the MPC kernel's placement / block-size knobs (environment variables
read at handle creation) over the BASELINE MPC configs.
Usage: python tools/tune_mpc.py [--scale F] [config ...]"""
import itertools
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
args = sys.argv[1:]
scale = "0.25"
if "--scale" in args:
    i = args.index("--scale")
    scale = args[i + 1]
    del args[i:i + 2]
configs = args or ["cfg3a_servo50", "cfg3b_dint50", "cfg4a_spacecraft100", "cfg4b_copoly100"]
for cfg in configs:
    big = cfg.startswith("cfg4")
    places = ["0", "1", "2", "3"]
    blocks = {"cfg4a_spacecraft100": ["32", "64"], "cfg4b_copoly100": ["64", "128"]}.get(cfg, ["32"])
    for place, block in itertools.product(places, blocks):
        env = dict(os.environ, FBSTAB_MPC_PLACE=place, FBSTAB_MPC_BLOCK=block)
        p = subprocess.run([sys.executable, os.path.join(HERE, "time_configs.py"), cfg,
                            "--scale", scale, "--reps", "2"], env=env, capture_output=True,
                           text=True, timeout=600)
        line = (p.stdout.strip().splitlines() or [p.stderr.strip()[-300:]])[-1]
        print(f"place={place} block={block} {line}", flush=True)
