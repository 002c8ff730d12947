import csv, collections, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; body=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ix['# Samples']]) for r in body)
print('total samples',tot,'instrs',len(body))
byop=collections.Counter(); bylong=collections.Counter(); cnt=collections.Counter(); ex=collections.Counter()
for r in body:
    s=r[ix['Source']].split()
    op=s[0] if not s[0].startswith('@') else s[1]
    op=op.split('.')[0]
    byop[op]+=int(r[ix['# Samples']]); bylong[op]+=int(r[ix['stall_long_sb']]); cnt[op]+=1; ex[op]+=int(r[ix['Instructions Executed']])
for op,v in byop.most_common(16): print("%-12s n=%4d exec %9d samples %6d (%.1f%%) long_sb %6d"%(op,cnt[op],ex[op],v,100*v/tot,bylong[op]))
print('--- top instrs')
top=sorted(body,key=lambda r:-int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 20]
for r in top: print(r[ix['# Samples']].rjust(6), 'long',r[ix['stall_long_sb']].rjust(5),'short',r[ix['stall_short_sb']].rjust(5),'wait',r[ix['stall_wait']].rjust(5), r[ix['Source']][:90])
