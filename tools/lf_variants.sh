#!/bin/bash
# times the loop filter of every variant library in vp8oclenc_b200/_variants/lf_*.so (kernel tuning aid, GPU only)
cd "$(dirname "$0")/.."
for so in vp8oclenc_b200/_variants/lf_*.so; do
  echo "== $so"
  VP8B200_ENGINE_LIB=$PWD/$so python - <<'PY'
import sys
sys.path.insert(0, "tools")
import lf_probe
for w, h in ((1920, 16), (1920, 1088)):
    print("  %5dx%-5d 3 planes %8.1f us   luma only %8.1f us   one chroma %8.1f us" % (w, h, lf_probe.probe(w, h), lf_probe.probe(w, h, luma_only=True), lf_probe.probe(w, h, luma_only=2)))
PY
done
