for k in node_time_d3_g2h5_ws_b3 node_time_d3_g2h5_ws_b4 node_time_d3_g2h5_b3; do
UDE_KERNEL=$k python bench.py --config c4 --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c4 $k value %.4e ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['kernel'][40:75])"
done
for k in node_d2_g2h5_ws_b3 node_d2_g2h5_ws_b4 node_d2_g2h5_b3; do
UDE_KERNEL=$k python bench.py --config c2 --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c2 $k value %.4e ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['kernel'][40:75])"
done
