"""Condenses the ncu launch lists of one training step (gpurun_out/launches_<model>.csv, written by
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--profile-from-start off --csv python tests/profile_step.py --model <model>`) into the tracked summaries:

  profiles/r02_launches_<model>.csv   per kernel: launches, total time, share of the step, DRAM bytes
  profiles/r02_traffic.json           DRAM traffic of the tensor-core engine kernels (bench.py's roofline.traffic)

ncu times are cold-cache and serialised: compare SHARES with bench.py's CUDA-event table, not absolutes."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENGINE = re.compile(r'gemm_sm100_kernel|attn_(fwd|bwd)_sm100_kernel')


def short(name):
    name = re.sub(r'saicv::(\(anonymous namespace\)|<unnamed>)::', '', name)
    name = re.sub(r'^void ', '', name)
    return name.split('(')[0].strip()


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    ii = hdr.index('ID')
    per = collections.defaultdict(dict)
    names = {}
    for r in rows[1:]:
        if len(r) <= vi or r[ii] == '':
            continue
        v = float(r[vi].replace(',', ''))
        unit = r[ui]
        if r[mi].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        elif r[mi].startswith('gpu__time'):
            v *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'ns ': 1e-3}.get(unit, 1)   # -> us
        per[r[ii]][r[mi]] = v
        names[r[ii]] = short(r[ki])
    return per, names


def main():
    traffic = {}
    for model in ('resnet50', 'vit_base_patch16'):
        src = os.path.join(ROOT, 'gpurun_out', f'launches_{model}.csv')
        if not os.path.exists(src):
            print('missing', src)
            continue
        per, names = load(src)
        agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for i, m in per.items():
            a = agg[names[i]]
            a[0] += 1
            a[1] += m.get('gpu__time_duration.sum', 0.0)
            a[2] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
        total = sum(a[1] for a in agg.values())
        out = os.path.join(ROOT, 'profiles', f'r02_launches_{model}.csv')
        with open(out, 'w') as f:
            f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off, '
                    f'ONE training step (tests/profile_step.py --model {model}, batch 256); cold-cache serialised times: compare shares\n')
            f.write('kernel,launches,total_us,share,dram_MB\n')
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f'{k},{a[0]},{a[1]:.1f},{a[1] / total:.4f},{a[2] / 1e6:.1f}\n')
        eng = [(names[i], m) for i, m in per.items() if ENGINE.search(names[i])]
        n = len(eng)
        dram = sum(m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0) for _, m in eng)
        traffic[model] = {'dram_bytes_per_launch': dram / max(n, 1), 'launches': n, 'dram_bytes_per_step': dram,
                          'note': ('sum of dram__bytes_read.sum + dram__bytes_write.sum over all gemm_sm100_kernel / attention launches of one '
                                   'training step (ncu, one pass per launch) divided by their count; compare with the algorithmic bytes per '
                                   'launch = roofline classes in the same JSON line')}
        print(model, 'kernels', len(agg), 'step us (ncu, serialised)', round(total), 'engine launches', n, 'engine DRAM GB', round(dram / 1e9, 2))
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w'), indent=1)


if __name__ == '__main__':
    sys.exit(main())
