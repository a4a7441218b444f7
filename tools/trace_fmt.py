import sys
lines=open(sys.argv[1]).read().split("==== RUN")[-1].strip().splitlines()
ev={}
for l in lines:
    if l.startswith("TRACE"):
        _,w,s,t=l.split(); ev[(w,int(s))]=float(t)
for j in [0,1,2,3,10,20,30,40,50,60]:
    g=lambda w: ev.get((w,j),float('nan'))
    prev=ev.get(("A.syrk",j-1), 0.0)
    print(f"j={j:2d} A: leaf_end {g('A.leaf'):8.1f} (leaf+gap {g('A.leaf')-prev:6.1f}) waited +{g('A.waited')-g('A.leaf'):6.1f} row +{g('A.row')-g('A.waited'):5.1f} syrk +{g('A.syrk')-g('A.row'):5.1f} | B: start {g('B.start'):8.1f} trsm +{g('B.trsm')-g('B.start'):5.1f} col +{g('B.col')-g('B.trsm'):5.1f} trail +{g('B.trail')-g('B.col'):6.1f} end {g('B.trail'):8.1f}")
