#!/bin/bash
# tools/collect_profiles_r02.sh -- turns what tools/gpu_profile_r02.sh (and the bench runs) left in gpurun_out/ into profiles/
R=r02
python tools/make_profile_md.py gpurun_out/r02_sqoa_encode.ncu-rep encode_block "SQOA encode, one launch = 16 cfg2 images (3840x2160 RGB), round 2" > profiles/${R}_sqoa_encode.md
python tools/make_profile_md.py gpurun_out/r02_qoi_encode.ncu-rep encode_block "QOI encode, one launch = 16 cfg2 images, round 2" > profiles/${R}_qoi_encode.md
python tools/make_profile_md.py gpurun_out/r02_sqoa_decode.ncu-rep sqoa_decode_kernel "SQOA decode, one launch = 16 cfg2 streams, round 2" > profiles/${R}_sqoa_decode.md
python tools/make_profile_md.py gpurun_out/r02_qoi_decode.ncu-rep qoi_rows "QOI decode (qoi_rows_kernel), one launch = 16 cfg2 streams, round 2" > profiles/${R}_qoi_decode_rows.md
cp gpurun_out/r02_bench_launches.csv profiles/${R}_bench_launches.csv
python - <<'PY'
import json,subprocess,csv
legs={}
def tobytes(v,unit): return float(v)*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[unit]
for leg in ("sqoa_encode","sqoa_decode","qoi_encode","qoi_decode"):
    out=subprocess.run(["ncu","-i",f"gpurun_out/r02_{leg}.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines())); h=rows[0]; u=rows[1]; r=rows[2]
    g=lambda n: r[h.index(n)]
    tot=tobytes(g("dram__bytes_read.sum"),u[h.index("dram__bytes_read.sum")])+tobytes(g("dram__bytes_write.sum"),u[h.index("dram__bytes_write.sum")])
    tunit=u[h.index("gpu__time_duration.sum")]
    t=float(g("gpu__time_duration.sum"))*{"us":1,"ms":1e3,"ns":1e-3,"s":1e6}.get(tunit,1)
    inst=float(g("smsp__inst_executed.sum"))
    legs[leg]={"dram_bytes_per_launch":int(tot/16),"dram_bytes_per_launch_of_16_images":int(tot),"images_per_launch":16,"ncu_time_us":round(t,1),
               "warp_instructions":int(inst),"thread_instructions_per_pixel":round(inst*32/(16*3840*2160),1),
               "issue_active_pct":round(float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")),1)}
d={"source":"ncu --set full --clock-control none captures of tools/prof_bench_legs.py (gpurun_out/r02_*.ncu-rep): one launch of the bench's own shape, a batch of 16 cfg2 images (398 MB of pixels, far larger than L2), so DRAM reads AND writes are visible; dram_bytes_per_launch is per IMAGE (bench.py multiplies by the images per launch)","legs":legs}
json.dump(d,open("profiles/r02_traffic.json","w"),indent=1)
print(json.dumps(legs,indent=1))
PY
for k in encode_block_kernelILi3ELb0 encode_block_kernelILi3ELb1 encode_block_kernelILi4ELb0 encode_block_kernelILi4ELb1; do
  ( echo "# tools/sass_mix.sh seqoia_b200/libsqoa_b200.so $k  (opcode mix of the kernel's SASS; UBLKCP = cp.async.bulk, SYNCS.* = mbarrier)"; bash tools/sass_mix.sh seqoia_b200/libsqoa_b200.so $k 300 ) > profiles/${R}_sass_mix_$k.txt
done
cp gpurun_out/r02_sqoabench.txt profiles/${R}_sqoabench.txt 2>/dev/null
