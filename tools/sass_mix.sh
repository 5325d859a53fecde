#!/bin/bash
# tools/sass_mix.sh <lib.so> <kernel-substring> -- instruction count and opcode mix of one kernel's SASS
cuobjdump -sass "$1" | awk -v k="$2" '/Function :/{on=index($0,k)>0} on' | grep -E "^\s+/\*[0-9a-f]{4,}\*/" \
  | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*$//; s/^@!?U?P[0-9T]+\s+//' > /tmp/sass_mix.txt
echo "instructions: $(wc -l < /tmp/sass_mix.txt)"
awk '{print $1}' /tmp/sass_mix.txt | sort | uniq -c | sort -rn | head -${3:-30}
