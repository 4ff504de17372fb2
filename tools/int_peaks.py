"""Measures the integer-multiply roofs the MSM / NTT rooflines refer to and writes gpurun_out/int_peaks.json
(copy it to profiles/int_peaks.json): jf_microbench(0) = independent IMAD.WIDE.U32 chains, jf_microbench(1) = dependent
256-bit Montgomery products in registers, five runs each, with the nvidia-smi clock record of the run."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_jellyfish_b200 as jf
import bench

ctx = jf.Context(0)
stop, rows = threading.Event(), []
th = threading.Thread(target=bench._clock_sampler, args=(stop, rows, 0), daemon=True)
th.start()
time.sleep(0.3)
imad, mont = [], []
t_end = time.time() + 3.0
while time.time() < t_end or len(imad) < 5:
    imad.append(ctx.microbench(0))
    mont.append(ctx.microbench(1))
stop.set(); th.join(timeout=2)
clocks = bench._clock_summary(rows)
name = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().splitlines()[0]
sm_mhz = clocks.get("sm_mhz") or 1965.0
out = {"gpu": name, "runs": len(imad),
       "imad_wide_u32_per_s": {"max": max(imad), "median": sorted(imad)[len(imad) // 2]},
       "mont_mul_256_per_s": {"max": max(mont), "median": sorted(mont)[len(mont) // 2]},
       "imad_wide_lanes_per_sm_clk": max(imad) / (148 * sm_mhz * 1e6),
       "limb_products_per_mont_mul_at_imad_rate": max(imad) / max(mont),
       "nominal_roof_148sm_x_32_lanes_x_clk": 148 * 32 * sm_mhz * 1e6,
       "clocks": clocks,
       "how": "tools/int_peaks.py: jf_microbench(0) / (1), repeated for 3 s; clocks sampled with nvidia-smi -lms 100 during the runs"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "int_peaks.json"), "w"), indent=1)
print(json.dumps(out))
