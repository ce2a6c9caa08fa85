import sys; sys.path.insert(0,'/root/repo')
exec(open('/root/repo/tools/layer_bench.py').read().replace("CONFIG1 = [", "CONFIG1 = [(0, 1, 3, 64, 512, 512, 4, 2, 2), (0, 2, 3, 64, 256, 256, 4, 2, 2), (0, 1, 4, 32, 128, 128, 3, 1, 1), (0, 1, 8, 64, 64, 64, 3, 1, 1),] + ["))
