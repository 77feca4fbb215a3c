import csv,collections,re,subprocess,os,tempfile,sys
rep=sys.argv[1]
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout.splitlines()
rows=list(csv.reader(src))
h=next(i for i,r in enumerate(rows) if r and r[0]=='Address')
hdr=rows[h]
ia=hdr.index('Address'); isrc=hdr.index('Source'); iex=hdr.index('Instructions Executed')
root='/root/repo'
lib=os.path.join(root,'srbd_horizon_b200/csrc/libsddp.so')
tmp=tempfile.mkdtemp()
subprocess.check_call(['cuobjdump','-xelf','all',lib],cwd=tmp,stdout=subprocess.DEVNULL)
cubin=[os.path.join(tmp,f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis=subprocess.run(['nvdisasm','-g','-c',cubin],capture_output=True,text=True).stdout.splitlines()
addr2line={};cur=None;inside=False
for ln in dis:
    if ln.startswith('\t.section\t.text.'): inside='solve_kernelI5SrbdTILb0EE9SmemSrbdTIS1_Lb0EELi4EE' in ln
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);',ln)
    if m: addr2line[int(m.group(1),16)]=cur
base=None
c=collections.Counter()
for r in rows[h+1:]:
    if len(r)<=iex: continue
    a=int(r[ia],16) if r[ia].startswith('0x') else int(r[ia])
    if base is None: base=a
    if re.search(r'\b(LDL|STL)\b',r[isrc]):
        c[(addr2line.get(a-base), 'LDL' if 'LDL' in r[isrc] else 'STL')]+=int(r[iex] or 0)
print('total local instr executed',sum(c.values()))
for k,v in c.most_common(30): print(v,k)
