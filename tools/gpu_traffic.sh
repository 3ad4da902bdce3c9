#!/bin/bash
# GPU box: DRAM traffic of every kernel in the pipeline's natural cache state (single-pass ncu, --cache-control none) and the
# launch list, for a given number of frames per launch.   usage: tools/gpu_traffic.sh <tag> <frames per launch>
TAG=${1:-r02af}; B=${2:-8}
mkdir -p gpurun_out
CMD="python bench.py --steps 16 --warmup 3 --pool 16 --batch $B --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:'k_tc|k_rmd' -s 12 -c 60 --csv \
  --log-file gpurun_out/${TAG}_traffic_warm_b$B.csv $CMD > gpurun_out/${TAG}_ncu_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tc|k_rmd' -s 12 -c 60 --csv \
  --log-file gpurun_out/${TAG}_launches_b$B.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
python - <<PY
import csv, collections, json
def load(path, metric_filter):
    d=collections.defaultdict(lambda: collections.defaultdict(list))
    for r in csv.reader(open(path)):
        if len(r)>5 and r[0].isdigit():
            name=r[4].split("(")[0]; metric=r[-3]; unit=r[-2]; val=float(r[-1].replace(",",""))
            scale={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9,"ns":1e-3,"nsecond":1e-3,"us":1,"usecond":1,"ms":1e3,"msecond":1e3}.get(unit,1)
            d[name][metric].append(val*scale)
    return d
t=load("gpurun_out/${TAG}_traffic_warm_b$B.csv",None)
out={}
tot=0
for k,v in t.items():
    rd=sum(v["dram__bytes_read.sum"])/max(1,len(v["dram__bytes_read.sum"])); wr=sum(v["dram__bytes_write.sum"])/max(1,len(v["dram__bytes_write.sum"]))
    out[k]={"read_MB":round(rd/1e6,2),"write_MB":round(wr/1e6,2),"launches":len(v["dram__bytes_read.sum"])}
    if k.startswith("k_tc") or "k_tc_fc" in k: tot+=rd+wr
l=load("gpurun_out/${TAG}_launches_b$B.csv",None)
for k,v in l.items():
    x=v["gpu__time_duration.sum"]; out.setdefault(k,{})["avg_us"]=round(sum(x)/len(x),1)
res={"frames_per_launch":$B,"cnn_dram_bytes_per_frame":tot/$B,"per_kernel_per_launch":out}
json.dump(res,open("gpurun_out/${TAG}_traffic_b$B.json","w"),indent=1)
print(json.dumps(res,indent=1))
PY
