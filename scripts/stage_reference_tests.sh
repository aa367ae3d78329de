#!/bin/bash
# Stage the reference's OWN test files beside the drop-in modules so that their
#   sys.path.insert(0, <tests>/../src)   and   runpy.run_path(<tests>/../src/<module>.py)
# resolve to THIS package's src/ (the reference's tests put their own src/ first, so running them from the reference
# tree never touches the drop-in).  Scratch only: ab_tmp/ is git-ignored and must not be committed.
#   build container:  bash scripts/stage_reference_tests.sh
#   GPU box:          cd ab_tmp/ref && python -m pytest tests -q -p no:cacheprovider
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
S=$ROOT/ab_tmp/ref
rm -rf "$S"; mkdir -p "$S/tests" "$S/stubs/matplotlib"
for t in test_fusion test_encoders test_attention test_uncertainty; do cp "$REF/tests/$t.py" "$S/tests/"; done
ln -s ../../multimodal-sensor-fusion-with-attention-rajeevatla_b200/src "$S/src"
# matplotlib is not in this image: the two tests that draw figures get a stand-in that records calls
cat > "$S/stubs/matplotlib/__init__.py" <<'PY'
from unittest.mock import MagicMock
import sys
def use(*a, **k): pass
def _save(path, *a, **k):            # a "figure" on disk, so that exists() checks of the tests hold
    with open(path, "wb") as f:
        f.write(b"stub figure")
def _fig():
    fig = MagicMock(name="fig")
    fig.savefig.side_effect = _save
    return fig
pyplot = MagicMock(name="matplotlib.pyplot")
pyplot.subplots.side_effect = lambda *a, **k: (_fig(), MagicMock(name="ax"))
pyplot.figure.side_effect = lambda *a, **k: _fig()
pyplot.savefig.side_effect = _save
sys.modules[__name__ + ".pyplot"] = pyplot
PY
cat > "$S/conftest.py" <<'PY'
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "stubs"))
PY
echo staged in $S
