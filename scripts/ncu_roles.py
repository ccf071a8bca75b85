#!/usr/bin/env python
"""Per-role stall breakdown of the warp-specialised tensor GEMM from an .ncu-rep with source.

    python scripts/ncu_roles.py gpurun_out/r01g_conv.ncu-rep [launch index]

Splits the SASS of gemm_bf16x3_kernel into its roles by marker instructions (LDTM = epilogue,
UTCHMMA = MMA issuer, UTMALDG = TMA producer, F2FP/LDG.128 = A loaders) and sums the warp-state
samples and executed instructions per role, plus the mbarrier try_wait spin loops by barrier
offset, which tell who waits for whom."""
import csv
import subprocess
import sys


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    skip = sys.argv[2] if len(sys.argv) > 2 else '0'
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:gemm_bf16x3',
                          '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:110])
    hdr = rows[1]
    ia, isrc, isamp, iex = (hdr.index(n) for n in ('Address', 'Source', '# Samples', 'Instructions Executed'))
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    seen, d = set(), []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[ia] == 'Address' or r[ia] in seen:
            continue
        seen.add(r[ia])
        d.append(r)
    d.sort(key=lambda r: int(r[ia], 16))
    marks = [(k, r[isrc]) for k, r in enumerate(d)]
    first = lambda pat, lo=0: next((k for k, s in marks if k >= lo and pat in s), len(d))
    k_epi = first('LDTM')
    k_mma = first('UTCHMMA')
    k_tma = first('UTMALDG')
    k_ldr = first('BAR.SYNC.DEFER_BLOCKING R', k_mma)
    k_exit = first('BAR.SYNC.DEFER_BLOCKING 0x0', k_tma)
    bounds = [('setup', 0, max(k_epi - 45, 0)), ('epilogue', max(k_epi - 45, 0), k_mma - 90), ('mma issuer', k_mma - 90, k_ldr - 200),
              ('A loaders', k_ldr - 200, k_tma - 20), ('tma producer', k_tma - 20, k_exit), ('teardown + spin loops', k_exit, len(d))]
    tot = sum(I(r[isamp]) for r in d)
    print('%d SASS instructions, %d warp samples' % (len(d), tot))
    for name, a, b in bounds:
        rs = d[a:b]
        n = sum(I(r[isamp]) for r in rs)
        ex = sum(I(r[iex]) for r in rs)
        st = {}
        for r in rs:
            for i in stall:
                st[hdr[i]] = st.get(hdr[i], 0) + I(r[i])
        top = ', '.join('%s %d' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print('  %-22s samples %6d (%4.1f%%)  warp-instructions %10d   %s' % (name, n, 100.0 * n / max(tot, 1), ex, top))
    print('mbarrier try_wait executions (spins) by barrier:')
    for k, r in enumerate(d):
        if 'SYNCS.PHASECHK' in r[isrc] and I(r[iex]) > 0:
            print('  %-70s executed %9d' % (r[isrc].strip()[:70], I(r[iex])))


if __name__ == '__main__':
    main()
