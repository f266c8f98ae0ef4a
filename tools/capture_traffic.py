"""profiles/dram_traffic.json from an ncu DRAM-byte launch list of ONE training step (bench.py quotes it as
`roofline.traffic` only while its `csrc_sha` matches the kernel sources).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/r2_dram_step.csv python tools/profile_step.py
    python tools/capture_traffic.py gpurun_out/r2_dram_step.csv profiles/dram_traffic.json
"""
import csv, hashlib, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha():
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'autoprog_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), 'rb') as fh:
            h.update(f.encode()); h.update(fh.read())
    return h.hexdigest()[:16]


rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
ki, mi, vi, ui, ii = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('Metric Unit'), H.index('ID')
per_launch = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui].lower()
    v *= {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1)
    per_launch.setdefault((r[ii], r[ki]), 0.0)
    per_launch[(r[ii], r[ki])] += v
groups = {'gemm_tc': r'gemm_tc2?_kernel', 'outlook': r'outlook_(fwd|bwd)_(mma|fma)_kernel', 'tlce': r'tlce_(fast_|cls_)?kernel',
          'mhsa': r'mhsa_(fwd|bwd|rowdot)_tc_kernel'}
per_op = {'tlce'}          # one op call per step = several launches: quote the step total as 'per launch'
out = {}
for g, pat in groups.items():
    vals = [v for (i, k), v in per_launch.items() if re.search(pat, k)]
    if vals:
        out[g] = {'launches': len(vals), 'dram_bytes_total': sum(vals), 'dram_bytes_per_launch': sum(vals) / (1 if g in per_op else len(vals))}
json.dump({'csrc_sha': sha(), 'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one training step (tools/profile_step.py: '
           'volo_d1, B=128, 224 px, bf16)', 'kernels': out}, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(out, indent=1))
