"""Per-source-line instruction counts of one kernel of an .ncu-rep (needs -lineinfo, --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top_n]   (exploration helper)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
acc = collections.OrderedDict()
total = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur, hdr = r[1].split('/')[-1], None
        continue
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and cur and len(r) == len(hdr):
        line, src = r[0], r[1]
        i_exec = hdr.index('Instructions Executed')
        i_samp = hdr.index('# Samples')
        try:
            n = int(r[i_exec]); smp = int(r[i_samp])
        except ValueError:
            continue
        key = (cur, int(line) if line.isdigit() else -1)
        e = acc.setdefault(key, [0, 0, src.strip()[:110], 0])
        e[0] += n; e[1] += smp; e[3] += 1
        total += n
tot_s = sum(e[1] for e in acc.values())
print('total warp instructions executed: %d   samples: %d' % (total, tot_s))
for (f, l), e in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.2f%% inst %5.2f%% samp  %3d sass  %-22s:%-4d %s' % (100.0 * e[0] / total, 100.0 * e[1] / max(tot_s, 1), e[3], f, l, e[2]))
