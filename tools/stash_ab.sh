# usage (under gpurun): bash tools/stash_ab.sh <tag>
# A/B of the reverse-sweep stash on C3: resident blocks per SM (= segments in flight = live stash bytes)
# and discard.global.L2 on / off.  For every variant: steps/s from a plain run, then DRAM bytes + L2 hit rate of ONE launch
# under ncu (metrics only, no clock control).  Output: gpurun_out/stash_ab_<tag>.txt
TAG=$1
OUT=gpurun_out/stash_ab_$TAG.txt
: > $OUT
for V in "mma2_h32_nt1" "mma2_h32_nt1:UDE_STASH_DISCARD=0" "mma2_h32_nt1:UDE_MAX_BLOCKS_PER_SM=2" "mma2_h32_nt1:UDE_MAX_BLOCKS_PER_SM=1" "mma2_h32_nt1:UDE_MAX_BLOCKS_PER_SM=1,UDE_STASH_DISCARD=0" "mma_h32"; do
  echo "=== $V" >> $OUT
  python tools/kernel_sweep.py --steps 10 "$V" 2>/dev/null | tail -1 | cut -c1-400 >> $OUT
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
      --clock-control none -k "regex:ude_mma2?_kernel" -s 3 -c 1 python tools/kernel_sweep.py --steps 1 "$V" 2>/dev/null | grep -E "dram__|gpu__time|lts__" >> $OUT
done
cat $OUT
