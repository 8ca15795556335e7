"""Write profiles/r<round>_ncu_traffic.json (DRAM bytes per launch, from `ncu --set full` captures) for bench.py's roofline.traffic.
usage: python tools/ncu_traffic.py <djpeg.ncu-rep> <conv.ncu-rep> [out.json]"""
import csv, io, json, subprocess, sys


def launches(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        v = float(r[i].replace(',', ''))
        u = units[i]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1)
    res = []
    for r in rows[2:]:
        res.append({'kernel': r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('<unnamed>::', ''),
                    'dram_read_bytes': val(r, 'dram__bytes_read.sum'), 'dram_write_bytes': val(r, 'dram__bytes_write.sum'),
                    'dram_bytes': val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum'), 'duration_us_under_ncu': val(r, 'gpu__time_duration.sum')})
    return res


if __name__ == '__main__':
    dj, cv = launches(sys.argv[1]), launches(sys.argv[2])
    out = {}
    f = [l for l in dj if 'fwd' in l['kernel']]
    b = [l for l in dj if 'bwd' in l['kernel']]
    if f:
        out['djpeg_fwd'] = dict(f[-1], workload='1280 x 128x128x3, q=50, soft', algorithmic_bytes=24 * 1280 * 128 * 128)
    if b:
        out['djpeg_bwd'] = dict(b[-1], workload='1280 x 128x128x3, q=50, soft', algorithmic_bytes=36 * 1280 * 128 * 128)
    if cv:
        top = max(cv, key=lambda l: l['duration_us_under_ncu'])
        out['conv_top_launch'] = dict(top, workload='FAN conv1 fprop: n1280 64x64 c32->64 k5', algorithmic_bytes=4 * 1280 * 64 * 64 * (32 + 64))
    path = sys.argv[3] if len(sys.argv) > 3 else 'profiles/r1_ncu_traffic.json'
    json.dump(out, open(path, 'w'), indent=1)
    print(json.dumps(out, indent=1))
