import csv,sys,re,json,subprocess
rep=sys.argv[1]; kern=sys.argv[2]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread ','launch__block_size','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum ','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed','smsp__average_warps_issue_stalled','smsp__issue_active.avg.pct','sm__icc_request_hit_rate','dram__bytes_read.sum ','dram__bytes_write.sum ','l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate']
for h,u,v in zip(hdr,units,vals):
    if any(w in h+' ' for w in want):
        try:
            if float(v.replace(',',''))<0.05 and 'stalled' in h: continue
        except: pass
        print(f"{h:86s} {v:>18s} {u}")
# per-function
elf=subprocess.run(['cuobjdump','-elf','/root/repo/roborugby_b200/librr_b200.so'],capture_output=True,text=True).stdout
syms=[]
for l in elf.splitlines():
    if kern not in l: continue
    p=l.split()
    if len(p)<7 or not p[0].startswith('0x'): continue
    try:
        v=int(p[1],16) if p[1].startswith('0x') else int(p[1]); s=int(p[2],16) if p[2].startswith('0x') else int(p[2])
    except: continue
    n=p[-1]
    if n.startswith('.text') or not n.startswith('$'): continue
    short=n.split('$')[2]
    m=re.match(r'_ZN2rr\d+([A-Za-z_0-9]+?)(I|E)', short); short=m.group(1) if m else short[:30]
    syms.append((v,s,short))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Address'); ii=hdr.index('Instructions Executed'); isn=hdr.index('# Samples'); it=hdr.index('Thread Instructions Executed')
cols={n:hdr.index(n) for n in ('stall_long_sb','stall_barrier','stall_wait','stall_selected','stall_branch_resolving','stall_short_sb','stall_no_inst','stall_math')}
base=int(data[0][ia],16)
agg={}
for r in data:
    off=int(r[ia],16)-base
    name='MAIN'
    for v,s,n in syms:
        if v<=off<v+s: name=n; break
    a=agg.setdefault(name,{'inst':0,'smp':0,'thr':0,**{c:0 for c in cols}})
    a['inst']+=int(r[ii] or 0); a['smp']+=int(r[isn] or 0); a['thr']+=int(r[it] or 0)
    for c,ix in cols.items(): a[c]+=int(r[ix] or 0)
ti=sum(a['inst'] for a in agg.values()); ts=sum(a['smp'] for a in agg.values())
print('total warp-inst',ti,'samples',ts)
print('%-28s %7s %7s %5s | %s'%('func','inst%','smp%','thr',' '.join(c[6:12] for c in cols)))
for n,a in sorted(agg.items(), key=lambda x:-x[1]['smp'])[:14]:
    print('%-28s %6.2f%% %6.2f%% %5.1f | %s'%(n[:28],100*a['inst']/ti,100*a['smp']/ts,a['thr']/max(a['inst'],1),' '.join('%6.2f'%(100*a[c]/ts) for c in cols)))
