"""Static SASS counts of the hot loops in libgfk.so (no GPU needed): instructions and FP64-pipe instructions per
candidate row of the persistent thinning kernel's update loop, registers per thread of every kernel, and the
Blackwell-specific opcodes present.  Writes profiles/r02_static_sass_counts.json, which bench.py reads to state the
ceiling the executed-vs-algorithmic FP64 instruction ratio puts on roofline.frac.

    python tools/sass_counts.py [path/to/libgfk.so]
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gradient_free_mcmc_postprocessing_b200', 'libgfk.so')
FP64 = ('DMUL', 'DADD', 'DFMA', 'DSETP', 'DMNMX')
ALGO_FLOPS = {'gf_d4': 8 * 4 + 16 + 3}


def functions(lib):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    cur, body = None, []
    for ln in out.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
        elif cur and re.match(r'\s+/\*[0-9a-f]{4}\*/', ln):
            op = re.sub(r'^\s*/\*[0-9a-f]+\*/\s*', '', ln)
            op = re.sub(r'^@!?U?P\d+\s+', '', op).strip()
            body.append(op.split()[0].rstrip(';') if op else '')
    if cur:
        yield cur, body


def row_periods(ops):
    """instruction windows between consecutive MUFU.RSQ64H (one per candidate row in the unrolled update loops)"""
    marks = [i for i, o in enumerate(ops) if o.startswith('MUFU.RSQ64H')]
    return [ops[a:b] for a, b in zip(marks, marks[1:])]


def registers(lib):
    out = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
    regs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r'Function (\S+):', ln)
        if m:
            cur = m.group(1)
        m = re.search(r'REG:(\d+)', ln)
        if m and cur:
            regs[cur] = int(m.group(1))
    return regs


def main():
    res = {'library': os.path.relpath(LIB, ROOT), 'how': 'cuobjdump -sass / -res-usage; a row period = the SASS between '
           'two consecutive MUFU.RSQ64H of an update loop; the median period is reported', 'kernels': {}}
    special = collections.Counter()
    for name, ops in functions(LIB):
        for o in ops:
            base = o.split('.')[0]
            if base in ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'UTCATOMSWS',
                        'SYNCS', 'LDGSTS', 'DMMA', 'HMMA', 'UTCBAR', 'UTCCP'):
                special[base] += 1
        if 'thin_persistent_kernel' not in name:
            continue
        periods = row_periods(ops)
        if not periods:
            continue
        by_len = sorted(periods, key=len)
        rep = by_len[(len(by_len) - 1) // 2]          # the median period: the prologue / epilogue windows are outliers
        common = len(rep)
        fp = collections.Counter(o.split('.')[0] for o in rep if o.split('.')[0] in FP64)
        res['kernels'][name] = {'instructions_per_row': common, 'fp64_pipe_per_row': sum(fp.values()), 'fp64_breakdown': dict(fp),
                                'row_periods_found': len(periods)}
    res['blackwell_opcodes'] = dict(special)
    regs = registers(LIB)
    res['registers'] = {k: v for k, v in sorted(regs.items()) if any(t in k for t in ('thin_persistent', 'ksd_rowsum', 'thin_step_chunked',
                                                                                     'kde_partial', 'dist_rowsum', 'ksd_mma', 'thin_step_bulk'))}

    def pick(strict, multi='Lb0E', wdd='Lb1E'):
        # thin_persistent_kernel<double, 4, STRICT, MULTI, WDD, TM=false>
        key = 'thin_persistent_kernelIdLi4E' + ('Lb1E' if strict else 'Lb0E') + multi + wdd + 'Lb0E'
        hits = [v for k, v in res['kernels'].items() if key in k]
        return hits[0] if hits else None

    flops = ALGO_FLOPS['gf_d4']
    summary = {}
    for label, strict in (('f64', True), ('f64-fast', False)):
        k = pick(strict)
        if k:
            summary[label] = {'fp64_pipe_instructions_per_pair': k['fp64_pipe_per_row'], 'instructions_per_pair': k['instructions_per_row'],
                              'algorithmic_flops_per_pair': flops,
                              # a fully busy FP64 pipe retires one instruction (<= 2 flops) per slot
                              'max_frac_at_full_fp64_pipe': flops / (2.0 * k['fp64_pipe_per_row']),
                              # one issue slot per instruction, one FP64 instruction per two slots
                              'max_frac_at_full_issue': flops / (2.0 * max(k['fp64_pipe_per_row'], k['instructions_per_row'] / 2.0))}
    res['c3_update_loop'] = summary
    out = os.path.join(ROOT, 'profiles', 'r02_static_sass_counts.json')
    json.dump(res, open(out, 'w'), indent=1, sort_keys=True)
    print(json.dumps(summary, indent=1))
    print('blackwell opcodes:', dict(special))
    print('wrote', out)


if __name__ == '__main__':
    main()
