"""cuobjdump -sass <binary> | python sassmix.py : opcode histogram per kernel (IMAD split by its .WIDE / .HI / .MOV ... forms),
used to check which pipe the rotates of rot_pipes.cu ended up on."""
import sys,re,collections
fn=None; cnt=collections.defaultdict(collections.Counter)
for l in sys.stdin:
    m=re.search(r'Function : (\S+)',l)
    if m: fn=m.group(1); continue
    m=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)',l)
    if m and fn:
        op=m.group(2); parts=op.split('.')
        cnt[fn][parts[0] + ('.'+parts[1] if parts[0]=='IMAD' and len(parts)>1 and parts[1] in('WIDE','HI','MOV','SHL','IADD','X') else '')]+=1
for fn in sorted(cnt): print(fn, dict(cnt[fn].most_common(14)))
