"""Per-level sweep timing with clusters of a forced size pulling tiles from the active-tile list (development aid)."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
for r in (0, 2, 3, 4, 6, 8, 12, 16):
    env = dict(os.environ)
    if r: env["VMORPH_R_DYN"] = str(r)
    p = subprocess.run([sys.executable, os.path.join(HERE, "time_levels.py"), cfg], env=env, stdout=subprocess.PIPE, text=True)
    print(f"=== R_DYN={r or 'default'}")
    print("\n".join(l for l in p.stdout.splitlines() if l.startswith("rep1")), flush=True)
