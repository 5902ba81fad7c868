"""Generates tests/golden/png_golden.npz: the synthetic PNG streams of tests/test_oracle_png.py::cases and what python-cv2 (here 4.13.0,
libpng) makes of them: cv2.imdecode(IMREAD_UNCHANGED) + cv2.cvtColor(RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY).  Run from the repo root."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_oracle_png as T   # noqa: E402

out = {}
for name, png in T.cases():
    out[name + "_png"] = np.frombuffer(png, np.uint8)
    for rgb in (True, False):
        out[f"{name}_gray{int(rgb)}"] = T.cv2_reference(png, rgb)
np.savez_compressed(T.GOLD, **out)
print(T.GOLD, T.GOLD.stat().st_size, "bytes,", len(out), "arrays")
