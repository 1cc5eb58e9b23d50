# Round-end evidence: launch list of one bench run + small ncu --set full captures (kept small: every
# captured kernel is replayed ~40 times and the report must stay far below 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?"
# first launch of a frame = split_im2col: find a frame start inside the list, derive the skip counts
F=$(python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); ii=h.index("ID")
starts=[int(r[ii]) for r in rows[1:] if 'split_im2col' in r[ik]]
print(starts[1])
PY
)
echo "frame starts at launch $F"
# ViT block 12 of that frame: LN, qkv, attention, proj, LN, fc1, fc2
timeout 900 ncu --set full --clock-control none --import-source on -s $((F + 5 + 84)) -c 7 -o gpurun_out/prof_vit_block $B > gpurun_out/ncu2.log 2>&1; echo "ncu vit exit $?"
# decoder + heads: the 54 launches after the ViT
timeout 1200 ncu --set full --clock-control none -s $((F + 180)) -c 54 --csv --page raw --log-file gpurun_out/ncu_decoder_raw.csv $B > gpurun_out/ncu4.log 2>&1; echo "ncu decoder exit $?"
ls -la gpurun_out | head -30; du -sh gpurun_out
