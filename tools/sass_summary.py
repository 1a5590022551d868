"""cuobjdump -sass ravnest_b200/ops/rvn_C.so | python tools/sass_summary.py > profiles/sass_mnemonics_r1.txt
Counts the instructions that prove the Blackwell path (tcgen05 / TMA / mbarrier) per kernel."""
import collections
import re
import subprocess
import sys

KEEP = ('UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTMAREDG', 'UTMACMDFLUSH', 'LDTM', 'STTM', 'SYNCS', 'UBLKCP', 'FENCE', 'MUFU', 'ATOMG',
        'RED', 'REDG', 'ATOM', 'LDG', 'STG', 'LDS', 'STS', 'HMMA', 'BAR', 'ELECT', 'UTCCP', 'UTCATOMSWS', 'ATOMS', 'MEMBAR', 'ERRBAR', 'CCTL')
cur, counts = None, collections.OrderedDict()
pat = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
for line in sys.stdin:
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.match(line)
    if m:
        base = m.group(1).split('.')[0]
        if base in KEEP:
            counts[cur][base] += 1
        counts[cur]['_total'] += 1
print('# SASS mnemonic counts per kernel in ravnest_b200/ops/rvn_C.so (cuobjdump -sass, sm_100a)')
print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit')
for k, c in counts.items():
    name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(.*', '', name)
    print('{:<64s} total={:<6d} '.format(name[:64], c['_total']) + ' '.join('{}={}'.format(a, b) for a, b in sorted(c.items()) if a != '_total'))
