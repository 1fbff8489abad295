import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from gdpathtracing_b200 import PathTracingCamera, scenes
sc = scenes.demo_scene(); grp = scenes.populate(sc)
cam = PathTracingCamera(); cam.fov=sc.fov; cam.geometry_group=grp; cam.denoising_mode=0
cam.set_window_size(1920,1080); cam.set_global_transform(sc.camera_transform12); cam.set_max_depth(8); cam.init()
for i in range(5): cam.render()
t=time.perf_counter()
for i in range(20): cam.render()
print('blocking ms/frame', (time.perf_counter()-t)/20*1e3, cam.stats()['k1_ms'])
tb=[];tw=[];k1=[]
t=time.perf_counter(); infl=0
for i in range(20):
    a=time.perf_counter(); cam.render_begin(); tb.append(time.perf_counter()-a); infl+=1
    if infl==2:
        a=time.perf_counter(); img,st=cam.render_wait(); tw.append(time.perf_counter()-a); k1.append(st['k1_ms']); infl-=1
while infl:
    img,st=cam.render_wait(); k1.append(st['k1_ms']); infl-=1
print('pipelined ms/frame', (time.perf_counter()-t)/20*1e3)
print('begin ms', np.round(np.array(tb)*1e3,3))
print('wait ms', np.round(np.array(tw)*1e3,3))
print('k1 ms', np.round(k1,3))
