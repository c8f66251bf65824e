"""Source lines of the flux kernel ranked by warp-stall samples, from an .ncu-rep captured with --import-source on
(and -lineinfo): usage  python profiles/scripts/ncu_hotlines.py <report.ncu-rep> [n_lines]
columns: share of samples, share of executed instructions, file, line, source, long-scoreboard / wait / no-instruction samples."""
import csv, collections, sys, subprocess
csv.field_size_limit(10**9)
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass,cuda'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
blocks=[];i=0
while i<len(rows):
    if rows[i] and rows[i][0]=='File Path':
        fp=rows[i][1]; hdr=rows[i+2]; j=i+3; data=[]
        while j<len(rows) and not (rows[j] and rows[j][0]=='File Path'):
            data.append(rows[j]); j+=1
        blocks.append((fp,hdr,data)); i=j
    else: i+=1
tot=0;toti=0;allL=[]
for fp,hdr,data in blocks:
    ln=hdr.index('Line No'); smp=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
    lsb=hdr.index('stall_long_sb'); w=hdr.index('stall_wait'); ni=hdr.index('stall_no_inst')
    for r in data:
        if len(r)>smp and r[ln]!='':
            try: s=int(r[smp] or 0); e=int(r[ie] or 0)
            except: continue
            tot+=s; toti+=e
            allL.append((s,e,fp.split('/')[-1],int(r[ln]),r[1].strip()[:70],r[lsb],r[w],r[ni]))
print('total samples',tot,'inst',toti)
allL.sort(reverse=True)
for x in allL[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print('%5.1f%% smp %5.1f%% inst'%(100*x[0]/tot,100*x[1]/toti), x[2][:14], x[3], x[4], 'lsb',x[5],'wait',x[6],'noinst',x[7])
