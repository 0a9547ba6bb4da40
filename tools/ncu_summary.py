"""Summarise ncu-rep files (read on the CPU box): python tools/ncu_summary.py <rep> [<rep> ...]"""
import csv, subprocess, sys, json
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','launch__shared_mem_config_size','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__sass_inst_executed_op_local_ld.sum','sm__sass_inst_executed_op_local_st.sum']
def summarise(rep):
    raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h:i for i,h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {'kernel': r[idx['Kernel Name']], 'grid': r[idx['Grid Size']], 'block': r[idx['Block Size']]}
        for k in KEYS:
            if k in idx: d[k] = (r[idx[k]] + ' ' + units[idx[k]]).strip()
        stall = [h for h in hdr if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
        d['top_stalls'] = sorted(((round(float(r[idx[h]].replace(',','')),2), h.split('stalled_')[1].split('_per')[0]) for h in stall), reverse=True)[:5]
        out.append(d)
    return out
if __name__ == '__main__':
    res = {rep: summarise(rep) for rep in sys.argv[1:]}
    print(json.dumps(res, indent=1))
