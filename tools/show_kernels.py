import json, sys, collections
k=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/kernels_slowfast.json'))
n=int(sys.argv[2]) if len(sys.argv)>2 else 30
tot=sum(x['ms'] for x in k)
print("total %.3f ms" % tot)
for x in sorted(k,key=lambda x:-x['ms'])[:n]:
    tf = x['flops']/(x['ms']*1e-3)/1e12 if x['ms']>0 else 0
    gb = x['bytes']/(x['ms']*1e-3)/1e9
    print("%-62s %-8s %8.1f us  %7.1f TF/s %8.1f GB/s  %6.2f GF %7.1f MB" % (x['name'][-62:], x['kind'], x['ms']*1e3, tf, gb, x['flops']/1e9, x['bytes']/1e6))
g=collections.defaultdict(float)
for x in k:
    nm=x['name']; g[nm.split('.res_blocks')[0] if 'res_blocks' in nm else nm]+=x['ms']
for kk,v in sorted(g.items(), key=lambda t:-t[1])[:14]: print("%-60s %.3f ms"%(kk,v))
