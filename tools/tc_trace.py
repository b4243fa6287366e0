"""Read the NSR_TC_TRACE dump of nsr_nerf_field_bwd_tc: [3 CTAs][16 wait codes][16 warps] cycle counters (lane 0 of every warp).
Row 0 = the warp's lifetime in the kernel; rows 1-9 = cycles spent waiting on: 1 producer<-stage empty, 2 MMA<-epilogue, 3 MMA<-stage full,
4 MMA<-DE drained, 5 epilogue<-MMA, 6 epilogue<-stage full, 7 epilogue<-all wgrad GEMMs, 8 scatter<-stage full, 9 scatter<-dE ready."""
import sys
t = [int(x) for x in open(sys.argv[1]).read().split()]
names = {1: 'producer<-empty', 2: 'mma<-epilogue', 3: 'mma<-stage', 4: 'mma<-DE drained', 5: 'epi<-mma', 6: 'epi<-stage', 7: 'epi<-wgrad', 8: 'scatter<-stage', 9: 'scatter<-dE'}
for c in range(3):
    blk = t[c * 256:(c + 1) * 256]
    life = blk[0:16]
    if not any(life):
        continue
    print(f'CTA sel {c}: lifetime (cycles) warp0 {life[0]}  warp1 {life[1]}  warp4 {life[4]}  warp8 {life[8]}  warp12 {life[12]}')
    for code in range(1, 10):
        row = blk[code * 16:(code + 1) * 16]
        if any(row):
            print(f'   {names[code]:18s}', ' '.join(f'w{w}:{100.0 * v / max(1, life[w]):.0f}%' for w, v in enumerate(row) if v))

# hand-off loop of CTA 0, tiles 2..: MMA lane stamps (committed, woke) per epilogue wait in [1024, 2048); epilogue lane stamps
# (acc ready, row stored, handed back) per step in [2048, 4096)
mma, epi = [x for x in t[1024:2048] if x], [x for x in t[2048:4096] if x]
n = min(len(mma) // 2, len(epi) // 3)
if n:
    import statistics as st
    legs = {'issue (woke -> group committed)': [], 'exec (committed -> acc visible)': [], 'epi work (ld, math, st.shared)': [],
            'epi fences + arrive': [], 'wake (arrive -> mma lane runs)': []}
    for i in range(n):
        committed, woke = mma[2 * i], mma[2 * i + 1]
        ready, stored, back = epi[3 * i], epi[3 * i + 1], epi[3 * i + 2]
        if i:
            legs['issue (woke -> group committed)'].append(committed - mma[2 * i - 1])
        legs['exec (committed -> acc visible)'].append(ready - committed)
        legs['epi work (ld, math, st.shared)'].append(stored - ready)
        legs['epi fences + arrive'].append(back - stored)
        legs['wake (arrive -> mma lane runs)'].append(woke - back)
    print('steps traced', n)
    for k, v in legs.items():
        by_step = [st.median(v[j::9]) for j in range(9) if v[j::9]]
        print(f'  {k:34s} median {st.median(v):7.0f}  per tile {sum(v) / (n / 9):8.0f}   by step {[int(x) for x in by_step]}')
