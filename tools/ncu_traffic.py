"""Extracts dram bytes per launch of one kernel from an `ncu --set full` report (raw page CSV) and records it in
profiles/pool_traffic.json under '<config>/<kernel>/<scenes>' -- what bench.py prints as roofline.traffic.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/rNN_x_raw.csv
    python tools/ncu_traffic.py profiles/rNN_x_raw.csv pool_tc32_kernel sgan_p 65536
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, kernel, config, scenes):
    rows = list(csv.reader(open(path)))
    head = rows[0]
    name_i = head.index('Kernel Name')
    rd, wr = head.index('dram__bytes_read.sum'), head.index('dram__bytes_write.sum')
    unit_rd, unit_wr = rows[1][rd], rows[1][wr]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    vals = [float(r[rd].replace(',', '')) * scale[unit_rd] + float(r[wr].replace(',', '')) * scale[unit_wr]
            for r in rows[2:] if kernel in r[name_i]]
    if not vals:
        raise SystemExit('no launch of %s in %s' % (kernel, path))
    out = os.path.join(ROOT, 'profiles', 'pool_traffic.json')
    table = json.load(open(out)) if os.path.isfile(out) else {}
    table['%s/%s/%d' % (config, kernel, int(scenes))] = {
        'dram_bytes_per_launch': sum(vals) / len(vals), 'launches': len(vals),
        'source': 'dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, %s' % os.path.relpath(path, ROOT)}
    json.dump(table, open(out, 'w'), indent=1, sort_keys=True)
    print(table)


if __name__ == '__main__':
    main(*sys.argv[1:5])
