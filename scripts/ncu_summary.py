"""Reads an `ncu --set full` report (.ncu-rep) and writes the per-kernel summary the bench
and DESIGN.md cite: duration, DRAM bytes, tensor-pipe activity, issue-slot use, registers.
  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_ncu_full_train_step.csv [profiles/r2_dram_traffic.json]
"""
import csv, io, json, subprocess, sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
           'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
           'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
           'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
           'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def main():
  rep, out_csv = sys.argv[1], sys.argv[2]
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr = rows[0]
  data = [r for r in rows[2:] if len(r) == len(hdr)]
  name_i = hdr.index('Kernel Name')
  cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
  with open(out_csv, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['id', 'kernel'] + [m for m, _ in cols])
    w.writerow(['', '(units)'] + [rows[1][i] for _, i in cols])
    for r in data:
      w.writerow([r[0], r[name_i][:70]] + [r[i] for _, i in cols])
  if len(sys.argv) > 3:
    def first(pattern):
      for r in data:
        if pattern in r[name_i]:
          return r
      return None
    def traffic(r):
      if r is None:
        return None
      rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', ''))
      wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', ''))
      units = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}      # one unit per COLUMN
      return (rd * units.get(rows[1][hdr.index('dram__bytes_read.sum')], 1) +
              wr * units.get(rows[1][hdr.index('dram__bytes_write.sum')], 1))
    out = {'forward': traffic(first('mlp_fwd_halves_kernel<2')), 'forward_nosave': traffic(first('mlp_fwd_halves_kernel<0')),
           'backward': traffic(first('mlp_bwd_wgrad_tmem_kernel')), 'info_nce': traffic(first('info_nce_kernel')),
           'first_layer': traffic(first('first_layer_grads_tiled_kernel')),
           'source': out_csv, 'unit': 'bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)'}
    if out['forward_nosave'] is None:     # the energy forward has its own capture: keep its entry
      try:
        prev = json.load(open(sys.argv[3]))
        for k in ('forward_nosave', 'forward_nosave_source'):
          if prev.get(k) is not None:
            out[k] = prev[k]
      except (OSError, ValueError):
        pass
    json.dump(out, open(sys.argv[3], 'w'), indent=1)
    print(out)


if __name__ == '__main__':
  main()
