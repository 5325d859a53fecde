#!/bin/bash
# tools/collect_profiles.sh -- turns the files a tools/gpu_profile_all.sh run left in gpurun_out/ into profiles/
R=${R:-r01}
for leg in sqoa_encode sqoa_decode qoi_encode; do
  case $leg in sqoa_encode) rx='encode_block'; t="SQOA encode, cfg2 (3840x2160 RGB)";; sqoa_decode) rx='sqoa_decode_kernel'; t="SQOA decode, cfg2";; qoi_encode) rx='encode_block'; t="QOI encode, cfg2";; esac
  python tools/make_profile_md.py gpurun_out/final_$leg.ncu-rep "$rx" "$t, round 1 final" > profiles/${R}_final_$leg.md
done
python tools/make_profile_md.py gpurun_out/final_qoi_decode.ncu-rep qoi_rows "QOI decode: qoi_rows_kernel (one launch per image), cfg2 (3840x2160 RGB), round 1 final" > profiles/${R}_final_qoi_decode_rows.md
(echo; echo "Instruction / stall-sample shares by code section (tools/ncu_sections.py):"; echo; echo '```'; python tools/ncu_sections.py gpurun_out/final_qoi_decode.ncu-rep; echo '```') >> profiles/${R}_final_qoi_decode_rows.md
cp gpurun_out/bench_launches.csv profiles/${R}_bench_launches.csv
python - <<'EOF'
import json,subprocess,csv
legs={}
def tobytes(v,unit): return float(v)*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[unit]
for leg in ("sqoa_encode","sqoa_decode","qoi_encode","qoi_decode"):
    out=subprocess.run(["ncu","-i",f"gpurun_out/final_{leg}.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines())); h=rows[0]; u=rows[1]
    tot=0; t=0; n=0
    for r in rows[2:]:
        tot+=tobytes(r[h.index("dram__bytes_read.sum")],u[h.index("dram__bytes_read.sum")])+tobytes(r[h.index("dram__bytes_write.sum")],u[h.index("dram__bytes_write.sum")])
        t+=float(r[h.index("gpu__time_duration.sum")]); n+=1
    legs[leg]={"dram_bytes_per_launch":int(tot),"kernels":n,"ncu_time_us":round(t,1)}
d={"source":"ncu --set full captures of tools/prof_legs.py on the cfg2 image (gpurun_out/final_*.ncu-rep, round 1, second half): dram__bytes_read.sum + dram__bytes_write.sum of the one kernel a leg launches (QOI decode: qoi_rows_kernel)","legs":legs}
json.dump(d,open("profiles/r01_traffic.json","w"),indent=1)
print(json.dumps(legs))
EOF
for f in bench_final cfg3_n1 cfg4_n1 cfg5_n1 reference_n1; do n=$f; [ $f = bench_final ] && n=bench_n1; [ $f = reference_n1 ] && n=reference_arm_n1; grep '^{' gpurun_out/$f.log | tail -1 > profiles/${R}_$n.json; done
