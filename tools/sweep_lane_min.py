import os, sys, json, subprocess
for kind, cfg in (("servo", "cfg3a_servo50"), ("dint", "cfg3b_dint50")):
    for scale in (0.03125, 0.125, 0.25, 0.5, 0.75):
        row = [kind, int(16384 * scale)]
        for lane in ("1", "0"):
            env = dict(os.environ, FBSTAB_MPC_LANE=lane)
            out = subprocess.run([sys.executable, "tools/time_configs.py", cfg, "--reps", "2", "--scale", str(scale)],
                                 capture_output=True, text=True, env=env).stdout.strip().splitlines()[-1]
            row.append(round(json.loads(out)["solves_per_s"]))
        print(row, flush=True)
