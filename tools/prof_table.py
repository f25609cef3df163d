import json, collections, sys
d=json.load(open(sys.argv[1]))
mb=d['micro_batch']
g=collections.OrderedDict()
for o in d['ops']:
    key=(o['kind'],o['Hout'],o['Cin'],o['Cout'],o['k'],o['stride'],o['residual'],o['block_n'])
    e=g.setdefault(key,[0,0.0,0.0]); e[0]+=1; e[1]+=o['ms']; e[2]+=o['gflop']
tot=sum(v[1] for v in g.values())
print("total ms",round(tot,3), "micro_batch", mb)
for k,v in sorted(g.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    tf = v[2]/v[1] if v[1]>0 else 0
    kind,H,Cin,Cout,kk,st,res,bn=k
    M=mb*H*H
    byts = M*Cout*2*(1+res) + (M*st*st)*Cin*2 if kind!='fc' else 0
    print(f"{kind:9s} H{H:4d} Cin{Cin:5d} Cout{Cout:5d} k{kk} s{st} res{res} bn{bn:4d} x{v[0]:3d}  {v[1]:7.3f} ms {100*v[1]/tot:5.1f}%  {1e3*v[1]/v[0]:7.1f} us/launch {tf:7.1f} TF/s  {byts*v[0]/v[1]/1e6:8.0f} GB/s")
