for sh in 4k3 4k4 big4 big3 1080p4; do
  timeout 200 python tools/time_legs.py --shape $sh --legs ${LEGS:-sqoa_encode} 2>&1 | grep -E "encode|decode|Error|error" | sed "s/^/$sh /"
done
