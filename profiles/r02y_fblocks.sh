set -u
OUT=gpurun_out; mkdir -p $OUT
for FB in 296 444 592 740; do
  echo -n "fblocks=$FB: "; TMD_SR_FBLOCKS=$FB timeout 200 python profiles/emulate_slabrun.py 8 40 2>&1 | tail -1
done
for FB in 592 740; do
  echo -n "world=4 fblocks=$FB: "; TMD_SR_FBLOCKS=$FB timeout 200 python profiles/emulate_slabrun.py 4 40 2>&1 | tail -1
done
