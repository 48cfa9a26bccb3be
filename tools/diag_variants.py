import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for lib in sorted((ROOT / "build_variants").glob("*.so")):
    env = dict(os.environ, MCSCE_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "diag_k2.py")], env=env, capture_output=True, text=True)
    print(lib.name); print("\n".join(l for l in r.stdout.splitlines() if "ratio-to-tol" in l), flush=True)
