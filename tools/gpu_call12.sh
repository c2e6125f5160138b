#!/bin/bash
mkdir -p gpurun_out
M=l1tex__data_pipe_tex_wavefronts.sum,l1tex__data_pipe_tex_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum,l1tex__t_requests_pipe_tex_mem_texture.sum,l1tex__t_sectors_pipe_tex_mem_texture.sum,l1tex__tex_writeback_active.sum.pct_of_peak_sustained_elapsed,l1tex__f_wavefronts.sum,l1tex__f_wavefronts.sum.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
for s in 1 9 11 13; do
timeout 300 ncu --metrics $M --clock-control none -k regex:k_fetch_quads -s $s -c 1 --csv --log-file gpurun_out/texb2_ncu_$s.csv tools/texbench/tex_bench2 > /dev/null 2>&1; echo "rc=$?"
done
cat gpurun_out/texb2_ncu_*.csv | grep -v "^==" | cut -d, -f5,13- 
