#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum
for la in 0 8; do
FL_PK_LOOKAHEAD=$la timeout 300 ncu --metrics $M -k regex:decode_persistent --clock-control none --csv --log-file $O/u_ncu_la$la.csv python tools/pk_phase_times.py > $O/u_ncu_la$la.log 2>&1
done
tail -20 $O/u_ncu_la0.csv | cut -c1-300; tail -20 $O/u_ncu_la8.csv | cut -c1-300
