import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
seen={}
for r in rows[hi+1:]:
    if len(r)>ix['Metric Value']:
        key=(r[ix['ID']], r[ix['Kernel Name']][:48])
        seen.setdefault(key,{})[r[ix['Metric Name']].split('.')[0].replace('smsp__','').replace('launch__','')]=r[ix['Metric Value']]
for k,v in list(seen.items())[-int(sys.argv[2]) if len(sys.argv)>2 else -4:]:
    print(k[1], v)
