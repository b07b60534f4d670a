import csv,re,sys
rows=list(csv.reader(open('/tmp/src.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=='Address')
H=rows[h]; body=[r for r in rows[h+1:] if len(r)==len(H)]
kern=sys.argv[1] if len(sys.argv)>1 else 'csp_batch_lean_kernelILi32ELb1'
lines=open('/tmp/sass_lines.txt').read().split('\n')
inside=False; cur=None; lmap=[]
for l in lines:
    if l.startswith('\t.section\t.text.'):
        inside = kern in l; continue
    if not inside: continue
    m=re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lmap.append(cur)
ix={n:H.index(n) for n in H}
tot=sum(int(r[ix['# Samples']] or 0) for r in body)
stalls=[n for n in H if n.startswith('stall_') and 'Not Issued' not in n]
agg={}
for n in stalls: agg[n]=sum(int(r[ix[n]] or 0) for r in body)
print('total samples',tot); print({k.replace('stall_',''):round(100*v/tot,1) for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:10]})
top=sorted(range(len(body)),key=lambda k:-int(body[k][ix['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for k in top:
    r=body[k]; s=int(r[ix['# Samples']] or 0)
    st=sorted(((int(r[ix[n]] or 0),n.replace('stall_','')) for n in stalls),reverse=True)[:2]
    print(f"{100*s/tot:5.2f}% {str(lmap[k] if k<len(lmap) else None):28s} {r[ix['Source']][:60]:60s} {st}")
