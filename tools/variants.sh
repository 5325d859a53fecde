for v in ${VARIANTS:-default}; do
for sh in ${SHAPES:-4k3 4k4 big4 1080p4}; do
  if [ "$v" = default ]; then L=""; else L=$PWD/gpurun_variants/libsqoa_b200_$v.so; fi
  SQOA_B200_LIB=$L timeout 200 python tools/time_legs.py --shape $sh --legs ${LEGS:-sqoa_encode} 2>&1 | grep -E "encode|decode|Error|error" | sed "s/^/$v $sh /"
done; done
