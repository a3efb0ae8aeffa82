"""Summarise an ncu report (read HERE with `ncu -i`): per kernel the headline metrics of the raw page and the warp-stall
samples of the source page grouped by SASS opcode.  Usage: python tools/ncu_summarize.py REPORT.ncu-rep OUT.json [note]"""
import collections, csv, io, json, re, subprocess, sys

RAW = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
       'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
       'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
       'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_active',
       'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'barrier', 'membar', 'mio_throttle',
          'lg_throttle', 'not_selected', 'dispatch_stall', 'no_instruction', 'branch_resolving']


def ncu(rep, page, extra=()):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ''
    rows = list(csv.reader(io.StringIO(ncu(rep, 'raw'))))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {'kernel': r[hdr.index('Kernel Name')], 'metrics': {}, 'stall_cycles_per_issue': {}}
        for m in RAW:
            if m in hdr and r[hdr.index(m)] != '':
                k['metrics'][m] = {'value': r[hdr.index(m)], 'unit': units[hdr.index(m)]}
        for s in STALLS:
            m = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if m in hdr and r[hdr.index(m)] != '':
                k['stall_cycles_per_issue'][s] = round(float(r[hdr.index(m)]), 3)
        kernels.append(k)
    text = ncu(rep, 'source', ('--print-source', 'sass')).split('\n')
    starts = [i for i, l in enumerate(text) if l.startswith('"Kernel Name"')] + [len(text)]
    seen = collections.Counter()
    for a, b in zip(starts[:-1], starts[1:]):
        name = next(csv.reader([text[a]]))[1]
        rd = list(csv.reader(text[a + 1:b]))
        if not rd or 'Source' not in rd[0] or '# Samples' not in rd[0]:
            continue
        h = rd[0]
        iS, iN = h.index('Source'), h.index('# Samples')
        cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
        cat, why, tot = collections.Counter(), collections.defaultdict(collections.Counter), 0
        for r in rd[1:]:
            if len(r) <= iN or not r[iN].isdigit():
                continue
            n = int(r[iN])
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
            op = m.group(2) if m else '?'
            tot += n
            cat[op] += n
            for c in cols:
                if r[c].isdigit() and int(r[c]):
                    why[op][h[c][6:]] += int(r[c])
        if not tot:
            continue
        # the source page lists every kernel once per view; keep the first with samples, match by order of appearance
        idx = [i for i, k in enumerate(kernels) if k['kernel'].split('(')[0].replace('ssn::', '') in name.replace('ssn::', '').replace('(int)', '').replace('(bool)', '')]
        tgt = kernels[idx[min(seen[name] // 1, len(idx) - 1)]] if idx else None
        seen[name] += 1
        if tgt is None or 'samples_by_opcode' in tgt:
            continue
        tgt['samples_total'] = tot
        tgt['samples_by_opcode'] = {op: {'pct': round(100.0 * n / tot, 1),
                                         'top_reasons': dict(why[op].most_common(3))} for op, n in cat.most_common(14)}
    json.dump({'report': rep, 'note': note, 'kernels': kernels}, open(out, 'w'), indent=1)
    for k in kernels:
        print(k['kernel'][:70], k['metrics'].get('gpu__time_duration.sum', {}).get('value'), list(k.get('samples_by_opcode', {}).items())[:3])


if __name__ == '__main__':
    main()
