"""Per-instruction stall summary of an .ncu-rep source page: segments between barriers + top stall sites."""
import csv, subprocess, sys
def main(path, topn=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:90])
    hdr = rows[1]; data = rows[2:]
    isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[isamp]) for r in data)
    print('total samples', tot, 'instr rows', len(data))
    agg = {}
    for r in data:
        for i, h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[i])
    print('stall totals', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
    cur = {'n': 0, 'samples': 0, 'exec': 0, 'start': 0}
    for k, r in enumerate(data):
        cur['n'] += 1; cur['samples'] += int(r[isamp]); cur['exec'] += int(r[iex])
        if 'BAR.SYNC' in r[isrc] or 'EXIT' in r[isrc]:
            cur['end'] = k; print('  segment', cur); cur = {'n': 0, 'samples': 0, 'exec': 0, 'start': k + 1}
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:topn]:
        st = {h[6:]: int(r[i]) for i, h in stall_cols if int(r[i]) > 0}
        print(f"{r[isamp]:>5} {r[iex]:>7} {r[isrc].strip()[:58]:58s}", dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))
if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
