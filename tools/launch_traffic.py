"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
one training step: per kernel name launches, device time and DRAM bytes; with --json also the per-family DRAM traffic
that bench.py reports as `roofline.traffic` (bytes per launch, cold caches, kernels serialised).
    python tools/launch_traffic.py gpurun_out/r02_launches.csv [--json profiles/r02_traffic.json]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ii, ki, mi, vi, ui = (hdr.index(c) for c in ('ID', 'Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    unit = r[ui]
    if r[mi] == 'gpu__time_duration.sum':
        v = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(unit, 1.0) * v
    else:
        v = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0) * v
    short = re.sub(r'\(CUtensor.*', '', r[ki])
    short = re.sub(r'\((const|int|float|long|__nv|void|unsigned).*', '', short)
    short = re.sub(r'^void ', '', short).replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    launches.setdefault(r[ii], {'name': short})[r[mi]] = v
seq = list(launches.values())
# exactly one step: from the first log-mel kernel to the launch before the next one
starts = [i for i, l in enumerate(seq) if l['name'].startswith('mel_logpower')]
if len(starts) >= 2:
    seq = seq[starts[0]: starts[1]]
elif starts:
    seq = seq[starts[0]:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for l in seq:
    a = agg[l['name']]
    a[0] += 1
    a[1] += l.get('gpu__time_duration.sum', 0.0)
    a[2] += l.get('dram__bytes_read.sum', 0.0)
    a[3] += l.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
print('one step: %d launches, %.1f us serialised (cold caches), DRAM read %.2f GB, write %.2f GB' % (
    len(seq), tot, sum(a[2] for a in agg.values()) / 1e9, sum(a[3] for a in agg.values()) / 1e9))
print('%10s %6s %5s %10s %10s  %s' % ('us', 'share', 'n', 'rd MB/l', 'wr MB/l', 'kernel'))
for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%10.1f %5.1f%% %5d %10.2f %10.2f  %s' % (t, 100 * t / tot, n, rd / n / 1e6, wr / n / 1e6, k[:100]))


def fam(pred):
    sel = [(n, rd + wr) for k, (n, t, rd, wr) in agg.items() if pred(k)]
    return sel


if '--json' in sys.argv:
    def per_call(pred, calls):
        tot_b = sum(b for _, b in fam(pred))
        return tot_b / calls if calls else None
    n_gn_f = sum(n for k, (n, *_r) in agg.items() if k.startswith('gn_fused_fwd'))
    n_gn_b = sum(n for k, (n, *_r) in agg.items() if k.startswith('gn_fused_bwd'))
    n_blk = sum(n for k, (n, *_r) in agg.items() if k.startswith('mqa_fwd'))
    out = {
        'source': 'ncu launch list of one step of `python bench.py --steps 2 --warmup 3 --no-cpu` (profiles/r02_ncu_launches.csv); '
                  'dram__bytes_read.sum + dram__bytes_write.sum per call, caches flushed between kernels',
        'gemm_family': sum(b for _, b in fam(lambda k: k.startswith('gemm_tc_kernel'))),
        'mel_forward': per_call(lambda k: k.startswith('mel_'), 1),
        'conv2_fwd': per_call(lambda k: k.startswith('conv_gemm_kernel<0'), 1),
        'conv2_dgrad': per_call(lambda k: k.startswith('conv_gemm_kernel<1'), 1),
        'conv2_wgrad': per_call(lambda k: k.startswith('conv_gemm_kernel<2'), 1),
        'groupnorm_fwd': per_call(lambda k: k.startswith('gn_fused_fwd'), n_gn_f),
        'groupnorm_bwd': per_call(lambda k: k.startswith('gn_fused_bwd'), n_gn_b),
        'dwconv31_fwd': per_call(lambda k: k.startswith('dwconv_fwd'), n_blk),
        'dwconv31_bwd': per_call(lambda k: k.startswith('dwconv_bwd'), n_blk),
        'mqa_attention_fwd': per_call(lambda k: k.startswith('mqa_fwd'), n_blk),
        'mqa_attention_bwd': per_call(lambda k: k.startswith(('mqa_bwd', 'attn_delta', 'attn_dq_finalize')), n_blk),
        'ctc_loss_fwd_bwd': per_call(lambda k: k.startswith('ctc_'), 1),
        'conv1_fwd': per_call(lambda k: k.startswith(('conv1_tc_fwd', 'conv1_fwd')), 1),
        'conv1_bwd': per_call(lambda k: k.startswith(('conv1_tc_bwd', 'conv1_bwd')), 1),
        'clip_adamw': per_call(lambda k: k.startswith(('sumsq_kernel', 'adamw_kernel')), 1),
    }
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)
    print('wrote', sys.argv[sys.argv.index('--json') + 1])
