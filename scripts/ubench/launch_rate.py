"""How many kernel launches per second does graph replay sustain on one GPU when D graphs of 39 tiny kernels on 5 streams
each (the shape of one pipelined forward) are in flight?  A floor for us-per-batch that no kernel tuning removes."""
import sys

import torch

D = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
cur = torch.cuda.current_stream(dev)
for per_side, label in ((9, "39 kernels: 4 on main, 4 x ~9 on side streams"), (4, "20 kernels"), (2, "12 kernels")):
    graphs, mains = [], []
    for d in range(D):
        main = torch.cuda.Stream(dev)
        sides = [torch.cuda.Stream(dev) for _ in range(4)]
        bufs = [torch.zeros(32, device=dev) for _ in range(5)]

        def body(main=main, sides=sides, bufs=bufs):
            m = torch.cuda.current_stream(dev)
            for li in range(4):
                bufs[0].add_(1.0)
                sides[li].wait_stream(m)
                with torch.cuda.stream(sides[li]):
                    for _ in range(per_side):
                        bufs[1 + li].add_(1.0)
            for s in sides:
                m.wait_stream(s)
        with torch.cuda.stream(main):
            body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            body()
        graphs.append(g)
        mains.append(main)

    def run(steps):
        for m in mains:
            m.wait_stream(cur)
        for i in range(steps):
            with torch.cuda.stream(mains[i % D]):
                graphs[i % D].replay()
        for m in mains:
            cur.wait_stream(m)
    run(4 * D)
    torch.cuda.synchronize()
    K = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    run(K)
    e1.record(cur)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / K
    nk = 4 + 4 * per_side
    print("%-48s %.1f us per replay, %.2f us per kernel (D = %d graphs in flight)" % (label, us, us / nk, D))
