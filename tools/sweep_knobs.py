"""Per-level sweep timing under different cluster sizes / CTA widths (development aid)."""
import os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
for cl, nw in ((1, "thr8"), (1, "thr16"), (1, "lat"), (2, "lat"), (4, "lat"), (8, "lat"), (16, "lat"), (8, "thr16"), (0, 0)):
    env = dict(os.environ)
    if cl: env["VMORPH_CLUSTER"] = str(cl)
    if nw: env["VMORPH_VARIANT"] = str(nw)
    r = subprocess.run([sys.executable, os.path.join(HERE, "time_levels.py"), cfg], env=env, stdout=subprocess.PIPE, text=True)
    print(f"=== cluster={cl or 'auto'} variant={nw or 'auto'}")
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("rep1")))
