"""Development probe: the bench's end-to-end figure alone (RIC_CHUNKS sweep)."""
import subprocess, sys, json, os
r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "..", "bench.py"), "--steps", "3", "--no-cpu-baseline",
                    "--configs3-images", "0", "--files-batch", "0", "--e2e-steps", "10"], capture_output=True, text=True)
l = json.loads(r.stdout.strip().splitlines()[-1])
print(os.environ.get("RIC_CHUNKS", "8"), "e2e", round(l["e2e"]["value"]), "seq", round(l["e2e"]["sequential_value"]), "roof", round(l["e2e"]["copy_only_value"]), "single", round(l["e2e"].get("single_call_value", 0)), l["e2e"].get("outputs_equal_across_slots"))
