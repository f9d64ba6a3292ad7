"""Turns the ncu outputs that tests/run_gpu_profile.sh left in gpurun_out/ into the small tracked
summaries under profiles/ (run in the build container; needs the `ncu` CLI to read .ncu-rep)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r01'


def short(name):
    name = re.sub(r'void |unnamed>::|<unnamed>::|\(anonymous namespace\)::', '', name)
    m = re.match(r'([\w:]+)(<[^(]*>)?\(', name)
    if m:
        tmpl = m.group(2) or ''
        return (m.group(1).split('::')[-1] + (tmpl if len(tmpl) < 24 else ''))[:64]
    return name[:64]


def launches(model):
    path = os.path.join(SRC, f'launches_{model}.csv')
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r['Metric Name'] != 'gpu__time_duration.sum':
            continue
        ns = float(r['Metric Value'].replace(',', ''))
        k = short(r['Kernel Name'])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(os.path.join(OUT, f'{TAG}_launches_{model}.csv'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off, ONE training step '
                f'(tests/profile_step.py --model {model}, batch 256); cold-cache serialised times: compare shares\n')
        f.write('kernel,launches,total_us,share\n')
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k},{n},{ns / 1e3:.1f},{ns / total:.4f}\n')
        f.write(f'TOTAL,{sum(a[0] for a in agg.values())},{total / 1e3:.1f},1.0\n')
    print(model, 'launch list:', len(rows), 'launches,', f'{total / 1e6:.2f} ms')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active']


def full(rep, out_name, note):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [w for w in WANT if w in idx]
    with open(os.path.join(OUT, out_name), 'w') as f:
        f.write(f'# {note}\n# ncu --set full --clock-control none --import-source on; units: ' +
                ', '.join(f'{c} [{units[idx[c]]}]' for c in cols) + '\n')
        f.write('launch,kernel,block,grid,' + ','.join(cols) + '\n')
        for i, r in enumerate(rows[2:]):
            f.write(f'{i},{short(r[idx["Kernel Name"]])},"{r[idx["Block Size"]]}","{r[idx["Grid Size"]]}",' +
                    ','.join(r[idx[c]].replace(',', '') for c in cols) + '\n')
    print(out_name, len(rows) - 2, 'launches')


def main():
    os.makedirs(OUT, exist_ok=True)
    for m in ('resnet50', 'vit_base_patch16'):
        launches(m)
    full('prof_gemm_r50.ncu-rep', f'{TAG}_ncu_gemm_resnet50_fwd.csv',
         'first 8 gemm_sm100_kernel launches of the ResNet-50 bs256 forward: stem GEMM (im2col matrix), layer1.0 conv1 1x1, conv2 3x3, conv3 1x1, downsample 1x1, layer1.1 ...')
    full('prof_gemm_vit.ncu-rep', f'{TAG}_ncu_gemm_vit_fwd.csv',
         'first 5 gemm_sm100_kernel launches of the ViT-B/16 bs256 forward: patch embedding, block0 qkv, proj, fc1, fc2')
    full('prof_attn_vit.ncu-rep', f'{TAG}_ncu_attention_vit.csv', 'attn_fwd_kernel / attn_bwd_kernel (mma.sync path) of ViT-B/16 bs256')
    full('prof_bnstats.ncu-rep', f'{TAG}_ncu_bn_colreduce_atomics_version.csv',
         'colreduce_kernel<0> (BN statistics) BEFORE the atomics were replaced by per-block partial rows (kept as the motivation)')


if __name__ == '__main__':
    main()
