"""How fast can the GPU front end retire graph kernel nodes?  n_streams graphs of n_nodes tiny kernels each."""
import sys, time, torch
n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 4600
for n_streams in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    xs = [torch.zeros(64, device='cuda') for _ in range(n_streams)]
    graphs = []
    for st, x in zip(streams, xs):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            x.add_(1.0)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=st):
                for _ in range(n_nodes):
                    x.add_(1.0)
        graphs.append(g)
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for _ in range(3):
            for st, g in zip(streams, graphs):
                with torch.cuda.stream(st):
                    g.replay()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    total = 3 * n_streams * n_nodes
    print('streams %2d: %.2f us per kernel node overall (%.0f k nodes/s), %.2f us per node per stream'
          % (n_streams, dt / total * 1e6, total / dt / 1e3, dt / (3 * n_nodes) * 1e6))
