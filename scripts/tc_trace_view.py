import sys
f=sys.argv[1]
ev=[l.split() for l in open(f) if l.startswith('TRACE')]
ev=[(int(a),int(b),int(c),int(d)) for _,a,b,c,d in ev]
n=len(ev)//2
ev=ev[n:]   # last launch
names={}
B_FULL=0;B_EMPTY=8;B_DFULL=16;B_DEMPTY=19;B_H1=22;B_H2EMPTY=29;B_YFULL=31;B_XFULL=32
def nm(role,tag):
    if tag>=200: return f"prod stage {tag-200} empty-wait"
    if tag>=100: return f"chunk {tag-100} done"
    if tag==99: return "tail issued"
    if tag==98: return "X published"
    if tag==B_FULL: return "FULL"
    if B_DFULL<=tag<B_DFULL+3: return f"DFULL{tag-B_DFULL}"
    if B_DEMPTY<=tag<B_DEMPTY+3: return f"DEMPTY{tag-B_DEMPTY}"
    if B_H1<=tag<B_H1+7: return f"H1_{tag-B_H1}"
    if B_H2EMPTY<=tag<B_H2EMPTY+2: return f"H2EMPTY{tag-B_H2EMPTY}"
    if tag==B_YFULL: return "YFULL"
    if tag==B_XFULL: return "XFULL"
    return str(tag)
base=min(e[2] for e in ev)
rows=sorted(ev,key=lambda e:e[2])
role=int(sys.argv[2]) if len(sys.argv)>2 else None
for r,tag,t0,t1 in rows:
    if role is not None and r!=role: continue
    print(f"{'  '*r*6}r{r} {t0-base:6d} +{t1-t0:5d} {nm(r,tag)}")
print("span",max(e[3] for e in ev)-base)
