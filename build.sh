#!/bin/bash
# convenience: build oracle + CUDA library (same as __graft_entry__.build())
cd "$(dirname "$0")" && python - "$@" <<'PY'
import sys, time, __graft_entry__ as g
t = time.time()
try:
    g.build_oracle(); g.build_product(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built in %.1fs" % (time.time() - t))
except Exception as e:
    print(str(e)[-6000:]); sys.exit(1)
PY
