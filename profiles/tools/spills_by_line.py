"""Register spills (STL / LDL in SASS) attributed to the outermost source line of a kernel file via nvdisasm -gi.
usage: spills_by_line.py build/X.o <kernel symbol substring> <file.cu> <min line>   (runs without a GPU)"""
import re,collections,sys,subprocess,os,glob,tempfile
obj, sym, fname, minline = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
d=tempfile.mkdtemp(); 
subprocess.run(['cuobjdump','-xelf','all',os.path.abspath(obj)],cwd=d,capture_output=True)
cub=glob.glob(d+'/*.cubin')[0]
txt=subprocess.run(['nvdisasm','-gi','-c',cub],capture_output=True,text=True).stdout
infn=False; cnt=collections.Counter(); ctx=[]; n=0
for l in txt.splitlines():
    if l.startswith('.text.'):
        infn = sym in l; continue
    if not infn: continue
    m=re.match(r'\s*//## File "(.*)", line (\d+)(.*)',l)
    if m:
        e=(m.group(1).split('/')[-1],int(m.group(2)))
        if 'inlined at' in l: ctx.append(e)
        else: ctx=[e]
        continue
    if re.match(r'\s*/\*[0-9a-f]{4,6}\*/',l):
        n+=1
        if ' STL' in l or ' LDL' in l:
            outer=[c for c in ctx if c[0]==fname and c[1]>minline]
            cnt[(outer[-1][1] if outer else -1, 'STL' if 'STL' in l else 'LDL')]+=1
print('instructions',n)
lines=open(os.path.join(os.path.dirname(os.path.abspath(obj)),'..','csrc',fname)).read().splitlines()
agg=collections.defaultdict(dict)
for (ln,k),v in cnt.items(): agg[ln][k]=v
for ln in sorted(agg): print(ln, agg[ln], (lines[ln-1].strip()[:100] if ln>0 else ''))
