"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list and the gradient step's shares.
usage: python profiles/launch_summary.py launches.csv "command line" > summary.txt"""
import csv, re, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
tot = collections.OrderedDict()
for r in rows:
    if r[0] == 'ID':
        hdr = r
        continue
    if hdr is None:
        continue
    name = r[hdr.index('Kernel Name')]
    val = float(r[hdr.index('Metric Value')])
    unit = r[hdr.index('Metric Unit')]
    ms = val * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
    k = re.sub(r'\(.*$', '', name)[:52]
    d = tot.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += ms
allms = sum(v[1] for v in tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print('per-kernel totals, cold-cache and serialised:')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%-54s n=%4d %10.3f ms  %5.1f%%  (%.3f ms per launch)' % (k, v[0], v[1], 100 * v[1] / allms, v[1] / v[0]))


def per(key):
    for kk, v in tot.items():
        if key in kk:
            return v[1] / v[0], v[0]
    return 0.0, 1


f, _ = per('forward_kernel_t<3')
w, nw = per('weights_kernel_t')
a, _ = per('apply_kernel<1, 1>')
srt = sum(v[1] for kk, v in tot.items() if 'RadixSort' in kk or 'pair_' in kk) / max(nw, 1)
step = f + w + a + srt
print()
print('gradient step under ncu (per call): forward %.2f + derivative walk %.2f + apply %.2f + pair sort/sums %.2f ms = %.2f ms; '
      'shares %.0f / %.0f / %.0f / %.0f %%' % (f, w, a, srt, step, 100 * f / step, 100 * w / step, 100 * a / step, 100 * srt / step))
