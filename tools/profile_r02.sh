#!/bin/bash
# Round-2 evidence run (one B200): the ncu launch list + DRAM bytes of one bench step and `ncu --set full` captures of the
# kernels DESIGN.md discusses, summarised ON THE BOX (the .ncu-rep files are too large to bring back).  Every command is
# bounded by `timeout`.  (compute-sanitizer is closed on this pool: profiles/r02_sanitizer_unavailable.txt.)
set -u
O=gpurun_out
NCU="ncu --clock-control none"
B="python bench.py --profile-only --steps 1 --warmup 1"
timeout 600 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -c 1500 --csv --log-file $O/traffic_r02.csv $B > $O/r02_ncu_list.out 2>&1; echo "launch list rc=$?"
cap() { # name, kernel regex, skip, count
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/r02_$1 $B > $O/r02_$1.out 2>&1; echo "$1 rc=$?"
  python profiles/summarize_ncu.py full /tmp/r02_$1.ncu-rep > $O/r02_$1_full.md 2>> $O/r02_$1.out
  python tools/ncu_hot_lines.py /tmp/r02_$1.ncu-rep 25 > $O/r02_$1_hot.txt 2>> $O/r02_$1.out
  rm -f /tmp/r02_$1.ncu-rep
}
cap text_gemms gemm_tcgen05 90 4        # layer 0 of the second step: qkv, attention-out, ffn1, ffn2 (folded LayerNorm variants)
cap l2_convs gemm_tcgen05 140 6         # layer2.0 conv2 / conv3+ds, layer2.1 conv1 / conv2 / conv3, layer2.2 conv1
cap l3_convs gemm_tcgen05 151 6         # layer3.0 conv1 / conv2 / conv3+ds, layer3.1 conv1 / conv2 / conv3
cap attention attention_short 12 1
cap bneck bneck64 3 3
cap stem_pre "stem_pool|preprocess_tiled" 2 2
ls -la $O | grep r02_ | tail -30
