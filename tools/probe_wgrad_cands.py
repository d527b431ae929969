"""wgrad time per candidate tiling (RSB_WGRAD_CAND) for the layers of the UNet: python tools/probe_wgrad_cands.py"""
import os, subprocess, sys
LAYERS = [(128, 96, 64), (128, 32, 32), (64, 64, 64), (64, 192, 128), (64, 32, 128), (32, 128, 128), (32, 384, 256), (32, 64, 256),
          (16, 256, 256), (16, 576, 512), (16, 128, 512), (8, 320, 320), (8, 256, 640)]
for (s, ci, co) in LAYERS:
    row = []
    for cand in ("", "0", "1", "2", "3", "4"):
        env = dict(os.environ)
        if cand:
            env["RSB_WGRAD_CAND"] = cand
        else:
            env.pop("RSB_WGRAD_CAND", None)
        r = subprocess.run([sys.executable, "tools/probe_wgrad_one.py", "2", str(s), str(s), str(s), str(ci), str(co)], env=env,
                           capture_output=True, text=True)
        out = [l for l in r.stdout.splitlines() if l.startswith("wgrad")]
        row.append(out[0].split(":")[1].split("ms")[0].strip() if out else "fail")
    print(f"{s}^3 {ci}x{co}: auto {row[0]} | A {row[1]} | B {row[2]} | C {row[3]} | D {row[4]} | E {row[5]}  (ms)", flush=True)
