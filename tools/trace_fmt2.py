import sys
lines=open(sys.argv[1]).read().split("==== RUN")[-1].strip().splitlines()
ev={}
for l in lines:
    if l.startswith("TRACE"):
        _,w,s,t=l.split(); ev[(w,int(s))]=float(t)
print("j : A.leaf_end  A.waited(+)  | B.start  trsm(+)  trail(+)  B.end | period")
for j in [0,1,2,3,5,10,15,20,25,30,40,50,60]:
    g=lambda w,jj=j: ev.get((w,jj),float('nan'))
    print(f"{j:2d}: {g('A.leaf'):9.1f} +{g('A.waited')-g('A.leaf'):6.1f} | {g('B.start'):9.1f} +{g('B.trsm')-g('B.start'):6.1f} +{g('B.trail')-g('B.trsm'):6.1f} {g('B.trail'):9.1f} | {g('A.leaf')-g('A.leaf',j-1):6.1f}")
