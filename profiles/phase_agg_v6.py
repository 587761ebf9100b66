import csv,sys,re
rows=list(csv.reader(open(sys.argv[1])))
cur=None;hdr=None;acc={};tot=0;stall={}
def phase(f,ln):
    if f=='device_common.cuh':
        if ln<=50: return 'hash murmur'
        if ln in (56,58,60): return 'stage (is_acgt/upcase/code)'
        if 66<=ln<=78: return 'rev2/revcomp'
        if 79<=ln<=105: return 'hash ascii expand'
        return 'misc dc'
    if f=='sketch_tile.cuh':
        if 100<=ln<=135: return 'stage_chunk'
        if 136<=ln<=150: return 'extract'
        if 151<=ln<=163: return 'any_bits (validity)'
        if 164<=ln<=210: return 'slow compare/hash_at'
        if 211<=ln<=258: return 'phase_canon (smem)'
        if 259<=ln<=295: return 'block minima/argmin (smem)'
        if 296<=ln<=310: return 'SeqModel validity/first'
        if 311<=ln<=322: return 'bounds'
        if 323<=ln<=375: return 'phase_runs (smem)'
        if 380<=ln<=420: return 'fast: window_minima'
        if 421<=ln<=452: return 'fast: canon roll'
        if 453<=ln<=468: return 'fast: prefix/suffix'
        if 469<=ln<=482: return 'fast: dispatch+valid loop'
        if 483<=ln<=504: return 'fast: run starts'
        if 505<=ln<=530: return 'fast: compaction'
        return 'tile misc'
    if f=='sketch_kernels.cu': return 'kernel:%d0s'%(ln//10)
    return f
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]
    elif len(r)>10 and r[0]=='Line No': hdr=r
    elif hdr and len(r)==len(hdr) and r[0] not in ('','Line No'):
        d=dict(zip(hdr[4:],r[4:]))
        try: n=float(d.get('Instructions Executed','0').replace(',','') or 0); st=float(d.get('Warp Stall Sampling (All Samples)','0').replace(',','') or 0)
        except: continue
        try: ln=int(r[0])
        except: continue
        ph=phase(cur,ln); acc[ph]=acc.get(ph,0)+n; stall[ph]=stall.get(ph,0)+st; tot+=n
ts=sum(stall.values())
for k,v in sorted(acc.items(),key=lambda x:-x[1])[:40]: print(f'{100*v/tot:6.2f}% inst {100*stall[k]/ts:6.2f}% stall  {k}')
