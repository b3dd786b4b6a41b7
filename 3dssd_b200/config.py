"""Backbone architecture tables, as data.

ARCH_3DSSD is the FIRST_STAGE.ARCHITECTURE list of /root/reference/configs/kitti/3dssd/3dssd.yaml:46-67, with the
16 positional fields documented at /root/reference/lib/core/config.py:207-219:
  0 xyz_index, 1 feature_index, 2 radius_list, 3 nsample_list, 4 mlp_list, 5 bn,
  6 fps_sample_range_list, 7 fps_method_list, 8 npoint_list, 9 former_fps_idx, 10 use_attention,
  11 layer_type, 12 scope, 13 dilated_group, 14 vote_ctr_index, 15 aggregation_channel
"""

ARCH_3DSSD = [
    [[0], [0], [0.2, 0.4, 0.8], [32, 32, 64], [[16, 16, 32], [16, 16, 32], [32, 32, 64]], True,
     [-1], ['D-FPS'], [4096],
     -1, False, 'SA_Layer', 'layer1', True, -1, 64],
    [[1], [1], [0.4, 0.8, 1.6], [32, 32, 64], [[64, 64, 128], [64, 64, 128], [64, 96, 128]], True,
     [-1], ['FS'], [512],
     -1, False, 'SA_Layer', 'layer2', True, -1, 128],
    [[2], [2], [1.6, 3.2, 4.8], [32, 32, 32], [[128, 128, 256], [128, 192, 256], [128, 256, 256]], True,
     [512, -1], ['F-FPS', 'D-FPS'], [256, 256],
     -1, False, 'SA_Layer', 'layer3', True, -1, 256],
    [[3], [3], [], [], [], True,
     [256, -1], ['F-FPS', 'D-FPS'], [256, 0],
     -1, False, 'SA_Layer', 'vote', False, -1, 256],
    [[4], [4], -1, -1, [128], True,
     [-1], [-1], [-1],
     -1, -1, 'Vote_Layer', 'vote', False, -1, -1],
    [[3], [3], [4.8, 6.4], [16, 32], [[256, 256, 512], [256, 512, 1024]], True,
     [-1], ['D-FPS'], [256],
     -1, False, 'SA_Layer', 'layer4', False, 5, 512],
]

# MODEL.MAX_TRANSLATE_RANGE (3dssd.yaml:39)
MAX_TRANSLATE_RANGE = (-3.0, -2.0, -3.0)
# MODEL.NETWORK.AGGREGATION_SA_FEATURE (3dssd.yaml:44)
AGGREGATION_SA_FEATURE = True
# points per scene after the loader's resampling (3dssd.yaml:36, lib/dataset/placeholders.py:26)
POINTS_NUM = 16384
INPUT_CHANNELS = 4  # x, y, z, intensity

# BASELINE.json configs[0] / SURVEY.md section 8d "config 1": one SA layer N=4096 -> 1024, plain ball query
ARCH_SINGLE_SA = [
    [[0], [0], [0.4], [32], [[64, 64, 128]], True,
     [-1], ['D-FPS'], [1024],
     -1, False, 'SA_Layer', 'layer1', False, -1, -1],
]


def layer_channels(arch, in_channels):
    """Feature channel count after every layer of `arch` (index 0 = network input)."""
    ch = [in_channels]
    for spec in arch:
        ltype, mlps, agg = spec[11], spec[4], spec[15]
        cin = ch[spec[1][0]]
        if ltype == 'SA_Layer':
            if len(spec[2]) == 0:
                ch.append(cin)
            elif agg is not None and agg != -1 and AGGREGATION_SA_FEATURE:
                ch.append(agg)
            else:
                ch.append(sum(m[-1] for m in mlps))
        elif ltype == 'Vote_Layer':
            ch.append(mlps[-1])
        elif ltype in ('SA_Layer_SSG_Last', 'FP_Layer'):
            ch.append(mlps[-1])
        else:
            raise ValueError(ltype)
    return ch
