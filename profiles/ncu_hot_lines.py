"""Per-source-line instruction and stall-sample shares from an .ncu-rep (needs --import-source on, -lineinfo).
usage: python profiles/ncu_hot_lines.py report.ncu-rep [top_n]"""
import sys, csv, subprocess, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; hdr = None; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-':     # per-line aggregate rows
        d = {}
        for k, v in zip(hdr, r):
            if k not in d: d[k] = v
        d['file'] = fname; data.append(d)
def num(x):
    try: return float(x)
    except Exception: return 0.0
tot = sum(num(d.get('Instructions Executed', 0)) for d in data)
tots = sum(num(d.get('# Samples', 0)) for d in data)
print('total warp instructions %.4g, stall samples %.4g' % (tot, tots))
data.sort(key=lambda d: -num(d.get('# Samples', 0)))
for d in data[:topn]:
    print('%5.1f%% inst %5.1f%% smp  %s:%s  %s' % (100 * num(d.get('Instructions Executed', 0)) / tot,
          100 * num(d.get('# Samples', 0)) / tots, d['file'], d['Line No'], d['Source'].strip()[:110]))
