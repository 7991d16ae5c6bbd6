"""dram__bytes_read.sum + dram__bytes_write.sum per kernel of ONE forward + backward chain, from an `ncu --set full`
report, into profiles/r02_ncu_traffic.json (what bench.py reports as roofline.traffic for the matching shape).

    python tools/ncu_traffic.py gpurun_out/rep.ncu-rep "T=32768,d_in=4096,d_out=4096,r=16,M=3,p=0.05" [profiles/r02_ncu_traffic.json]
"""
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            'profiles', 'r02_ncu_traffic.json')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

    def val(r, name):
        return float(r[idx[name]].replace(',', '')) * scale.get(units[idx[name]], 1.0)

    per = {}
    order = []
    for r in rows[2:]:
        if len(r) <= idx['Kernel Name']:
            continue
        name = r[idx['Kernel Name']].split('(')[0].replace('moka::', '').replace('void ', '')
        rec = dict(dram_read=val(r, 'dram__bytes_read.sum'), dram_write=val(r, 'dram__bytes_write.sum'),
                   duration_us=float(r[idx['gpu__time_duration.sum']].replace(',', '')))
        per.setdefault(name, []).append(rec)
        if name not in order:
            order.append(name)
    # the LAST captured launch of every kernel (buffers rotate, so nothing is warm in L2 that would not be in a step)
    chain = {n: per[n][-1] for n in order}
    total = sum(v['dram_read'] + v['dram_write'] for v in chain.values())
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[key] = dict(source=os.path.basename(rep), traffic_bytes=total, kernels=chain,
                     note='dram__bytes_read.sum + dram__bytes_write.sum of one moka_forward + moka_backward chain (ncu --set full, '
                          '--clock-control none); dirty lines still in the 126 MB L2 at kernel end are not counted')
    json.dump(data, open(out, 'w'), indent=1, sort_keys=True)
    print(key, 'traffic', total / 1e6, 'MB')


if __name__ == '__main__':
    main()
