"""join ncu's per-SASS-instruction counts with nvdisasm line info: python sass_lines.py rep kernel_regex cubin func_substr"""
import csv,sys,subprocess,re,collections
rep,kre,cubin,fsub=sys.argv[1:5]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--kernel-name','regex:'+kre],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# may hold several launches: take the first kernel block
blocks=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur=[];blocks.append(cur);continue
    if cur is not None: cur.append(r)
b=blocks[0]; hdr=b[0]; ie=hdr.index('Instructions Executed'); st=hdr.index('Warp Stall Sampling (All Samples)')
ins=[(int(r[0],16),r[1].strip(),int(r[ie] or 0),int(r[st] or 0)) for r in b[1:] if len(r)>ie and r[0].startswith('0x')]
base=ins[0][0]
dis=subprocess.run(['nvdisasm','-c','-g',cubin],capture_output=True,text=True).stdout
infn=False; line=None; m_off={}
for l in dis.splitlines():
    if l.startswith('\t.section') or l.startswith('.section'): infn = fsub in l
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m: line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: m_off[int(m.group(1),16)]=line
agg=collections.Counter(); sagg=collections.Counter(); tot=0
for a,t,n,s in ins:
    ln=m_off.get(a-base); agg[ln]+=n; sagg[ln]+=s; tot+=n
print('total warp instr',tot, 'static',len(ins))
srcs={}
for (ln,n) in agg.most_common(int(sys.argv[5]) if len(sys.argv)>5 else 40):
    txt=''
    if ln:
        f=[p for p in ('/root/repo/legged_games_gym_b200/csrc/'+ln[0],) ]
        try:
            srcs.setdefault(ln[0],open(f[0]).read().splitlines()); txt=srcs[ln[0]][ln[1]-1].strip()[:100]
        except Exception: pass
    print(f"{n:>10} {100*n/tot:5.1f}% stall{sagg[ln]:>6} {ln} {txt}")
