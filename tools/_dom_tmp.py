import json, os, shutil, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.path.join(REPO, "earl_benchmark_b200")
main = os.path.join(pkg, "libearl_b200.so")
shutil.copy(main, main + ".orig")
kit = """
import bench, torch, json
out = bench.run_kitchen(torch.device('cuda:0'), 0, 1, 1965.0, 148, False)
print('RES kitchen', json.dumps({'value': round(out['value']), 'redone': out['work']['redone_states']}), flush=True)
"""
for name, code in (("orig", kit), ("kdom4", kit)):
    src = main + ".orig" if name == "orig" else os.path.join(pkg, "build", "variants", f"lib_{name}.so")
    shutil.copy(src, main)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=REPO), capture_output=True, text=True, cwd=REPO)
    res = [l for l in r.stdout.splitlines() if l.startswith("RES")]
    print(name, " | ".join(res) if res else r.stderr[-400:], flush=True)
shutil.copy(main + ".orig", main)
