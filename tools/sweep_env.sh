# usage: bash tools/sweep_env.sh  -- A/B of environment switches of libzvx.so on the bench workload (developer tool)
run() { name=$1; shift; env "$@" python tools/launch_table.py > gpurun_out/sw_$name.txt 2>&1; python - "$name" <<'PY'
import sys,collections
name=sys.argv[1]
d=collections.OrderedDict()
for line in open(f'gpurun_out/sw_{name}.txt'):
    f=line.split()
    if len(f)<6 or not f[0].isdigit(): continue
    k=f"{f[1]}:{f[2]}"; d[k]=d.get(k,0)+float(f[f.index('ms=')+1])
print(name, ' '.join(f"{k.replace(':st=','')}={v:.3f}" for k,v in d.items() if 'conv' in k), 'total=%.3f'%sum(d.values()))
PY
}
run base A=1
run smem56 ZVX_CONV_SMEM_KB=56
run smem72 ZVX_CONV_SMEM_KB=72
run base2 A=1
run smem56b ZVX_CONV_SMEM_KB=56
run smem72b ZVX_CONV_SMEM_KB=72
