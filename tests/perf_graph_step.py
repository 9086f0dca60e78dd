"""Not a test: eager vs CUDA-graphed training step (unopticalflow_b200.train.GraphedTrainStep) at B=8 and B=2."""
import sys, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import unopticalflow_b200 as u
from unopticalflow_b200 import train as T
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.benchmark = True
for B in (8, 2):
    torch.manual_seed(0)
    m1 = u.Model_flow(T.KITTI_CFG).cuda(); m2 = u.Model_flow(T.KITTI_CFG).cuda(); m2.load_state_dict(m1.state_dict())
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    xs = [torch.rand(B, 3, 768, 832, device='cuda') for _ in range(4)]
    opt = T.make_optimizer(m1)
    g = T.GraphedTrainStep(m2, xs[0], w)
    # the graphed model has taken `warmup`+1 steps on xs[0]; bring the eager one to the same state
    for _ in range(4): T.train_step(m1, opt, xs[0], w)
    for i in range(3):
        le = float(T.train_step(m1, opt, xs[i % 4], w)); lg = float(g(xs[i % 4]))
        print('B=%d step %d eager %.6f graphed %.6f' % (B, i, le, lg))
    def timeit(f, n=20):
        for i in range(3): f(i)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): f(i)
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    print('B=%d eager %.2f ms  graphed %.2f ms' % (B, timeit(lambda i: T.train_step(m1, opt, xs[i % 4], w)), timeit(lambda i: g(xs[i % 4]))))
