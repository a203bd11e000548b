#!/bin/bash
# BASELINE.json configs[4]: box2d1r 40960x40960 and box3d1r 1024^3 slab-decomposed over 1/2/4/8 B200 (strong scaling,
# global grid fixed), plus the headline 1d2r job (weak scaling) -- one box, back to back.  Usage (8-GPU box):
#   bash profiles/run_config_e.sh gpurun_out/scale   ->  gpurun_out/scale_<workload>_n<N>.json
out=${1:-gpurun_out/scale}
run() {  # name N bench-args...
    name=$1; n=$2; shift 2
    if [ "$n" = 1 ]; then
        python bench.py --gpus 1 --steps 3 --warmup 3 "$@" > ${out}_${name}_n1.json 2> ${out}_${name}_n1.err
    else
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
            bench.py --gpus $n --steps 3 --warmup 3 "$@" > ${out}_${name}_n$n.json 2> ${out}_${name}_n$n.err
    fi
    echo "$name n=$n rc=$? $(python -c "import json,sys; d=json.load(open('${out}_${name}_n$n.json')); print(round(d['value'],1), 'GStencil/s', round(d['ms_per_step'],2), 'ms/step', d['config']['decomposition'][:40])" 2>&1 | tail -1)"
}
ngpu=$(nvidia-smi -L | wc -l)
for n in ${NS:-8 4 2 1}; do
    [ $n -le $ngpu ] || continue
    run box2d1r_40960 $n --shape box2d1r --dims 40960,40960 --times 100 --scaling strong
    run box3d1r_1024 $n --shape box3d1r --dims 1024,1024,1024 --times 100 --scaling strong
    run 1d2r_weak $n --no-shapes --no-cpu
done
