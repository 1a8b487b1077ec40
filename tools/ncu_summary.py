"""profiles/r02_summary tables from an `ncu --page raw --csv` export of tools/profile_step.py (one eager step):
    python tools/ncu_summary.py profiles/r02_ncu_step_raw.csv profiles/r02_profile_step_ops.txt [frames_per_step]
Prints per-group and per-launch duration, DRAM bytes, tensor-pipe activity, achieved TFLOP/s; writes
profiles/traffic.json (DRAM bytes per frame per kernel group, read by bench.py's roofline.traffic)."""
import csv
import json
import os
import sys

raw, opsf = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 32
rows = list(csv.reader(open(raw)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
units = rows[1]                                               # row 1 = one (scaled) unit per column
data = [r for r in rows[2:] if len(r) == len(hdr)]
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6,          # times -> us
         'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}   # sizes -> bytes


def num(r, name):
  try:
    return float(r[col[name]].replace(',', '')) * SCALE.get(units[col[name]], 1.0)
  except Exception:
    return float('nan')


ops = []
seen = False
for l in open(opsf):
  if l.strip() == 'OPS':
    seen = True
    continue
  if seen and l.split() and l.split()[0].isdigit():
    ops.append(l.split()[1:])
kern = [r for r in data]
out = []
j = 0
for r in kern:
  name = r[col['Kernel Name']].split('(')[0]
  out.append({'kernel': name, 'grid': r[col['Grid Size']] if 'Grid Size' in col else '',
              'us': num(r, 'gpu__time_duration.sum'),
              'dram_rd': num(r, 'dram__bytes_read.sum'), 'dram_wr': num(r, 'dram__bytes_write.sum'),
              'tensor_pct': num(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') if
              'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed' in col else float('nan'),
              'issue_pct': num(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active') if
              'sm__inst_issued.avg.pct_of_peak_sustained_active' in col else float('nan'),
              'lsu_pct': num(r, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed') if
              'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed' in col else float('nan')})
# label: the conv launches appear in plan order; non-conv kernels by name
conv_ops = [o for o in ops if o[0] == 'conv']
ci = 0
for o in out:
  o['op'] = ''
  if any(k in o['kernel'] for k in ('conv_tc_kernel', 'dcn_persist_kernel', 'conv_halo_kernel', 'conv_simt_kernel')):
    if ci < len(conv_ops):
      o['op'] = conv_ops[ci][1]
      o['a_mode'] = conv_ops[ci][3] if len(conv_ops[ci]) > 3 else ''
      ci += 1
groups = {}
for o in out:
  k = o['kernel']
  # kernel FUNCTIONS: dcn_persist_kernel (window DCN, C_out <= 128), conv_tc_kernel split into its DCN launches
  # (C_out = 256 / non-window) and its plain gather launches, conv_halo_kernel
  g = 'dcn_persist' if 'dcn_persist' in k else 'dcn_tc' if ('conv_tc' in k and o.get('a_mode') in ('1', '2')) else \
      'conv_tc_plain' if 'conv_tc' in k else \
      'conv_halo' if 'conv_halo' in k else 'decode' if 'decode' in k else 'track' if 'track_step' in k else \
      'upsample' if 'upsample' in k else 'other'
  o['group'] = g
  gr = groups.setdefault(g, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0, 'tw': 0.0})
  gr['n'] += 1; gr['us'] += o['us']; gr['rd'] += o['dram_rd']; gr['wr'] += o['dram_wr']
  gr['tw'] += o['us'] * (o['tensor_pct'] if o['tensor_pct'] == o['tensor_pct'] else 0.0)
tot = sum(g['us'] for g in groups.values())
print('one eager step, %d frames: %d launches, %.1f us of kernels (ncu: serialised, cold caches -- compare shares)\n' % (frames, len(out), tot))
print('| group | launches | us | share | DRAM read MB | DRAM written MB | tensor pipe active % (time-weighted) |\n|---|---|---|---|---|---|---|')
for k, g in sorted(groups.items(), key=lambda kv: -kv[1]['us']):
  print('| %s | %d | %.1f | %.1f %% | %.1f | %.1f | %.1f |' % (k, g['n'], g['us'], 100 * g['us'] / tot, g['rd'] / 1e6, g['wr'] / 1e6,
                                                             g['tw'] / max(g['us'], 1e-9)))
print('\n| # | op | kernel | grid | us | DRAM rd MB | DRAM wr MB | tensor % | issue % | L1/smem data pipe % |\n|---|---|---|---|---|---|---|---|---|---|')
for i, o in enumerate(out):
  print('| %d | %s | %s | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |' % (i, o['op'], o['kernel'][-28:], o['grid'], o['us'], o['dram_rd'] / 1e6,
                                                                            o['dram_wr'] / 1e6, o['tensor_pct'], o['issue_pct'], o['lsu_pct']))
tj = os.path.join(os.path.dirname(os.path.abspath(raw)), 'traffic.json')
cur = json.load(open(tj)) if os.path.exists(tj) else {}
cfg = os.environ.get('CT_CFG', 'coco_tracking')
cur[cfg] = {k: (g['rd'] + g['wr']) / frames for k, g in groups.items()
            if k in ('dcn_persist', 'dcn_tc', 'conv_tc_plain', 'conv_halo', 'decode', 'upsample')}
json.dump(cur, open(tj, 'w'), indent=1, sort_keys=True)
