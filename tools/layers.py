"""Print a bench.py --layer-report JSON as a table: python tools/layers.py gpurun_out/layers.json [filter] [top]"""
import json, sys
d = json.load(open(sys.argv[1]))
flt = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
steps, rows = d['steps'], d['layers']
tot = sum(r['ms'] for r in rows) / steps
print('total ms/step %.2f' % tot)
for r in [r for r in rows if flt in r['key']][:top]:
    print('%-45s calls %3d  ms/step %7.3f  %5.1f%%  tflops %s' % (r['key'], r['calls'] / steps, r['ms'] / steps, 100 * r['ms'] / steps / tot, ('%.1f' % r['tflops']) if r['tflops'] else '-'))
