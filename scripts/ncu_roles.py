#!/usr/bin/env python
"""Stall reasons and instruction share of the inflate kernel by ROLE (source-line ranges of bgzf_inflate_tps.cuh) from an .ncu-rep
captured with --import-source on.  usage: ncu_roles.py rep name:lo-hi [name:lo-hi ...]   (e.g. decoder:871-1010 copy:405-590)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
cur=None; hdr=None; kern=0; agg=collections.defaultdict(lambda: collections.Counter()); lines={}
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': fn=r[1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0]!='':
        try: ln=int(r[0])
        except: continue
        if 'inflate_tps' not in fn: continue
        d={}
        for k,v in zip(hdr,r): d.setdefault(k,v)
        key=(cur,ln)
        for k,v in d.items():
            if k.startswith('stall_') and 'Not Issued' not in k:
                try: agg[key][k]+=int(v)
                except: pass
        try:
            agg[key]['inst']+=int(d['Instructions Executed']); agg[key]['smp']+=int(d['# Samples'])
        except: pass
        lines[key]=d['Source'].strip()[:80]
# roles by line range in bgzf_inflate_tps.cuh (ranges passed as args: name:lo-hi)
roles=[a.split(':') for a in sys.argv[2:]]
tot=collections.Counter(); per=collections.defaultdict(collections.Counter)
for (f,ln),c in agg.items():
    role='other:'+f
    if f=='bgzf_inflate_tps.cuh':
        role='tps-unassigned'
        for name,rng in roles:
            lo,hi=map(int,rng.split('-'))
            if lo<=ln<=hi: role=name; break
    per[role].update(c); tot.update(c)
print('total inst',tot['inst'],'samples',tot['smp'])
for role,c in sorted(per.items(), key=lambda kv:-kv[1]['smp']):
    st={k[6:]:v for k,v in c.items() if k.startswith('stall_') and v}
    s=sum(st.values()) or 1
    print('%-28s inst %5.1f%% smp %5.1f%% | '%(role,100*c['inst']/tot['inst'],100*c['smp']/tot['smp'])+' '.join('%s %.0f%%'%(k,100*v/s) for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:7]))
