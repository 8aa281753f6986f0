import sys, numpy as np
for l in open(sys.argv[1]):
    if not l.startswith('timeline thread'): continue
    name, vals = l.split(':', 1)
    v = np.array([int(x) for x in vals.split()])
    if 'thread 0' in name:
        g = v[:len(v) // 6 * 6].reshape(-1, 6)
        print("thread 0, per group: barrier | issue | commit+prefetch | wait | epilogue ; total", v[-1])
        for i in range(len(g) - 1):
            print("%2d  %5d %5d %5d %5d %5d" % (i, g[i,1]-g[i,0], g[i,2]-g[i,1], g[i,3]-g[i,2], g[i,5]-g[i,3], g[i+1,0]-g[i,5]))
    else:
        g = v[:len(v) // 4 * 4].reshape(-1, 4)
        print("thread 32, per group: barrier | ->wait | wait | epilogue ; total", v[-1])
        for i in range(len(g) - 1):
            print("%2d  %5d %5d %5d %5d" % (i, g[i,1]-g[i,0], g[i,2]-g[i,1], g[i,3]-g[i,2], g[i+1,0]-g[i,3]))
