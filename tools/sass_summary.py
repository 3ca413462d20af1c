"""Per-kernel SASS opcode summary of the in-tree library (the evidence table of B200_PROFILING.md: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, HMMA = legacy tensor path).

    python tools/sass_summary.py [lfm_quant_b200/_lfmq.so] > profiles/r02_sass_summary.txt
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'lfm_quant_b200/_lfmq.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
WATCH = ('UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UTMASTG', 'UBLKPF', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS', 'SYNCS',
         'HMMA', 'HGMMA', 'MUFU', 'STG', 'LDG', 'STL', 'LDL', 'ACQBULK', 'CCTL', 'ERRBAR', 'MEMBAR')
fn, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        for w in WATCH:
            if op.startswith(w):
                key = w + ('.MULTICAST' if 'MULTICAST' in op else '')
                counts[fn][key] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print('# SASS opcode counts per kernel of %s (cuobjdump -sass); columns = mnemonics of interest' % so)
for name, pretty in zip(counts, demangle):
    c = counts[name]
    if not total[name]:
        continue
    short = re.sub(r'\(.*', '', pretty)
    print('%-70s instr %6d  %s' % (short[:70], total[name], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
