import os, sys, tempfile, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from prospect_public_b200.run import run
from prospect_public_b200.toy import annealing_yaml
tmp = tempfile.mkdtemp()
paths = bench.toy_inputs(tmp)[0]
pkw = dict(bench.workload_profile_kw('profile30', 1)[1])
run(annealing_yaml(paths, os.path.join(tmp, 'warm'), jobtype='profile', write=True, **pkw))
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable()
run(annealing_yaml(paths, os.path.join(tmp, 'ttp'), jobtype='profile', write=True, **pkw))
pr.disable(); print('wall', time.perf_counter() - t0)
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
