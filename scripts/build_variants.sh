#!/bin/bash
# Builds libfst_b200.so and the tuning variants under libfst_b200/variants/ in parallel (see profiles/README.md).
cd "$(dirname "$0")/.."
mkdir -p libfst_b200/variants
b() { python -c "
from libfst_b200 import build as b
import sys
b.build(force=True, defines=[d for d in sys.argv[2:]], out=(None if sys.argv[1]=='main' else 'libfst_b200/variants/'+sys.argv[1]+'.so'))
print(sys.argv[1], 'ok')
" "$@" 2>&1 | tail -5; }
b main &
for spec in "$@"; do   # name:DEF1,DEF2
  name=${spec%%:*}; defs=${spec#*:}
  b $name ${defs//,/ } &
done
wait
ls -la libfst_b200/*.so libfst_b200/variants/
