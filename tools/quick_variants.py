# quick kernel timing of the lattice update variants (device-resident, fluid + cells as in bench)
import os, sys, json, subprocess
for env in sys.argv[1:]:
    e = dict(os.environ); e["HCG_MOMENT_ONLY"] = "1"
    for kv in env.split(","):
        if kv:
            k, v = kv.split("="); e[k] = v
    r = subprocess.run([sys.executable, "bench.py", "--steps", "30", "--warmup", "4", "--no-cpu-baseline", "--no-parity-check"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(env, "ms/step", round(d["ms_per_step"], 4), {k.replace("kernel:", ""): round(v, 4) for k, v in d["kernel_ms_per_step"].items()}, flush=True)
    except Exception as ex:
        print(env, "FAILED", r.stderr[-500:], flush=True)
