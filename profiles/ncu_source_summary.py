import csv, sys, re
path=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(path)))
# multiple kernels: split on "Kernel Name" rows
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'hdr':None,'rows':[]}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    cur['rows'].append(r)
which=int(sys.argv[3]) if len(sys.argv)>3 else -1
for bi,b in enumerate(blocks):
    if which>=0 and bi!=which: continue
    h={k:i for i,k in enumerate(b['hdr'])}
    tot=sum(int(r[h['# Samples']]) for r in b['rows'])
    tinst=sum(int(r[h['Instructions Executed']]) for r in b['rows'])
    print('==',bi,b['name'][:100],'samples',tot,'inst',tinst,'n_sass',len(b['rows']))
    # stall totals
    st=[k for k in b['hdr'] if k.startswith('stall_') and 'Not Issued' not in k]
    agg={k:sum(int(r[h[k]]) for r in b['rows']) for k in st}
    print('  stalls:',' '.join('%s=%.1f%%'%(k[6:],100*v/max(tot,1)) for k,v in sorted(agg.items(),key=lambda kv:-kv[1]) if v))
    # opcode histogram by executed
    ops={}
    for r in b['rows']:
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)',r[h['Source']]);
        op=m.group(2).split('.')[0] if m else '?'
        ops[op]=ops.get(op,0)+int(r[h['Instructions Executed']])
    print('  ops:',' '.join('%s=%.1f%%'%(k,100*v/tinst) for k,v in sorted(ops.items(),key=lambda kv:-kv[1])[:18]))
    top=sorted(range(len(b['rows'])),key=lambda i:-int(b['rows'][i][h['# Samples']]))[:topn]
    for i in sorted(top):
        r=b['rows'][i]
        ss=' '.join('%s=%s'%(k[6:],r[h[k]]) for k in st if int(r[h[k]])>0)
        print('  %5d %-70s smp=%5s ex=%8s %s'%(i,r[h['Source']].strip()[:70],r[h['# Samples']],r[h['Instructions Executed']],ss))
