"""Post-filter for the LGR_EMU_TSAN=2 run (CTAs of a launch left unordered): drop the reports that are only the reuse of the
emulated __shared__ storage by consecutive CTAs (location = a static of libemu, or the dynamic shared buffer allocated in
cuda_runtime.h) and print what is left -- races between CTAs on real global buffers.  Usage: python tsan_filter.py <log_path prefix>"""
import re, sys, glob
txt=''.join(open(f).read() for f in glob.glob(sys.argv[1]+'.*'))
reports=txt.split('==================')
keep=[]
tot=0
for r in reports:
    if 'WARNING: ThreadSanitizer' not in r: continue
    tot+=1
    loc=re.search(r'Location is (.*)', r)
    loc=loc.group(1) if loc else '?'
    if loc.startswith('global') : continue          # emulated __shared__ static (or emulator state)
    if 'heap block' in loc and 'cuda_runtime.h' in r.split('Location is')[1]: continue   # dynamic shared buffer / fiber stacks
    keep.append(r)
print('reports', tot, 'after filtering shared-storage reuse', len(keep))
for r in keep[:6]:
    lines=[l for l in r.split('\n') if re.search(r'WARNING|#0 |#1 |Location|of size', l)]
    print('\n'.join(l[:170] for l in lines[:10])); print('--')
