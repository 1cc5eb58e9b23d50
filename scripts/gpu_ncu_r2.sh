# Round-2 evidence (gpurun --timeout 3000 -- bash scripts/gpu_ncu_r2.sh; then scripts/ncu_summarise.py --tag r2): launch list of one bench run + small ncu --set full captures (kept small: every
# captured kernel is replayed ~40 times and the report must stay far below 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
# (eager launches on one stream: ncu serialises kernels anyway, and the launch list then reads in program order)
export DEPTHPRO_GRAPH=0 DEPTHPRO_STREAMS=0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-video"
timeout 600 $B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
# weight upload + LayerNorm folding take ~1270 launches before the first frame; the IDs in the list are
# relative to the first captured launch
SKIP=1400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 600 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?"
# frame structure from the kernel names: a frame starts at split_im2col; ViT block 12 = qkv, attention, proj,
# fc1, fc2 around the frame's 13th attention launch; the decoder follows the frame's last layernorm launch
read F V D N <<< $(SKIP=$SKIP python - <<'PY'
import csv, os
skip = int(os.environ["SKIP"])
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); ii=h.index("ID")
ids=[(int(r[ii]), r[ik]) for r in rows[1:]]
starts=[i for i,(_,k) in enumerate(ids) if 'split_im2col' in k]
a,b=starts[0],starts[1]
fr=ids[a:b]
att=[i for i,(_,k) in enumerate(fr) if 'attention_tc' in k]
ln=[i for i,(_,k) in enumerate(fr) if 'layernorm_kernel' in k]
print(skip+fr[0][0], skip+fr[att[12]-1][0], skip+fr[ln[-1]+1][0], len(fr)-(ln[-1]+1))
PY
)
echo "frame starts at launch $F, ViT block 12 at $V, decoder at $D ($N launches)"
timeout 900 ncu --set full --clock-control none --import-source on -s $V -c 5 -o gpurun_out/prof_vit_block $B > gpurun_out/ncu2.log 2>&1; echo "ncu vit exit $?"
timeout 600 ncu -i gpurun_out/prof_vit_block.ncu-rep --page raw --csv > gpurun_out/ncu_vit_block_raw.csv 2> gpurun_out/ncu3.log; echo "ncu vit export exit $?"
timeout 1200 ncu --set full --clock-control none -s $D -c $N --csv --page raw --log-file gpurun_out/ncu_decoder_raw.csv $B > gpurun_out/ncu4.log 2>&1; echo "ncu decoder exit $?"
ls -la gpurun_out | head -40; du -sh gpurun_out
