"""config-4 per-process mode (14 processes x B = 1 shim on one GPU): python tools/procs_probe.py <label>"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

if __name__ == "__main__":
    r = bench.run_rollout_processes(0, n_steps=40)
    print(sys.argv[1] if len(sys.argv) > 1 else "", json.dumps({k: v for k, v in r.items() if k not in ("workload", "reference_cpu_ms_per_step")}))
