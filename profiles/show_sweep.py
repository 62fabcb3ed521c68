import json,sys
txt=open(sys.argv[1]).read()
dec=json.JSONDecoder(); i=0; objs=[]
while i < len(txt):
    while i < len(txt) and txt[i] != '{': i+=1
    if i>=len(txt): break
    try:
        o,j=dec.raw_decode(txt,i); objs.append(o); i=j
    except Exception as e:
        i+=1
sh=lambda t:{k.replace(' done','').replace('barrier after ','B').replace('reduce','r').replace(' start','s').replace(' end','e'):v for k,v in t.items()}
full=len(sys.argv)>2
for d in objs:
    if 'stages' in d:
        print(d['stages'],d['push_blocks'],d['stage_fracs'],'STEP',d['step_ms'],'xchg',d['exchange_only_ms'],'red',d['reduce_only_ms'],d['stage_edges_rank0'])
        if d['timeline_ms_rank0']: print('   r0',sh(d['timeline_ms_rank0']))
    elif full:
        print('   r%d'%d['rank'],sh(d['timeline_ms']))
