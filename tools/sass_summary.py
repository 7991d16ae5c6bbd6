"""Per-kernel counts of the Blackwell instructions that prove the tcgen05 / TMA / TMEM path, from `cuobjdump -sass`
of the built libmoka.so -> profiles/rNN_sass_summary.md.      python tools/sass_summary.py profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'moka_b200', 'libmoka.so')
MNEMONICS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'ELECT', 'R2UR', 'HMMA', 'MUFU.EX2', 'RED.E', 'ACQBULK',
             'UTCATOMSWS', 'UBLKCP', 'FENCE.VIEW.ASYNC']


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_sass_summary.md')
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', name).replace('moka::', '')
            cur = re.sub(r'^void ', '', cur)
            kernels[cur] = collections.Counter()
            kernels[cur]['instructions'] = 0
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]['instructions'] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                kernels[cur][mn] += 1
    digest = open(os.path.join(ROOT, 'moka_b200', '.libmoka.stamp')).read().strip()[:16]
    tot = collections.Counter()
    lines = ['# r02: SASS summary of moka_b200/libmoka.so (sm_100a)', '',
             f'`cuobjdump -sass moka_b200/libmoka.so`, build stamp `{digest}` (sha256 of csrc/ + include/moka.h + flags).',
             'UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / .st (TMEM <-> registers; k_down_ts stages its A operand with STTM and its UTCHMMA reads `tmem[..], gdesc[..]`: the TS form), UTMALDG / UTMASTG = TMA tile',
             'load / store, SYNCS = mbarrier ops, ELECT = elect.sync (one-lane issue of the whole-warp MMA / TMA loops, round 2),',
             'R2UR = vector -> uniform register moves feeding those instructions, HMMA = warp-level mma.sync (attention kernels).', '',
             '| kernel | SASS instr | ' + ' | '.join(MNEMONICS[:11]) + ' |', '|---|---|' + '---|' * 11]
    for k, c in kernels.items():
        if not any(c[m] for m in MNEMONICS[:6]) and not c['HMMA']:
            continue
        lines.append(f'| `{k[:70]}` | {c["instructions"]} | ' + ' | '.join(str(c[m]) for m in MNEMONICS[:11]) + ' |')
        tot.update(c)
    lines.append('| **total (kernels listed)** | ' + str(tot['instructions']) + ' | ' + ' | '.join(str(tot[m]) for m in MNEMONICS[:11]) + ' |')
    lines += ['', f'Kernels in the library: {len(kernels)}; the rest (warp-FMA path, metadata builders, AdamW, RoPE, keep-mask hook) use none of the above.']
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[-8:]))


if __name__ == '__main__':
    main()
