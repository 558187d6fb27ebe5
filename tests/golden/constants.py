"""Constants of the reference configuration used by the fixtures (configure/cfgs.py:36-44, traincfg.yaml:56)."""
PART_LIST = ['head', 'neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'left_feet', 'right_ham',
             'right_shank', 'right_feet', 'left_arm', 'left_forearm', 'left_hand', 'right_arm', 'right_forearm',
             'right_hand']
NOLEAF = ['neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'right_ham', 'right_shank', 'left_arm',
          'left_forearm', 'right_arm', 'right_forearm']
MEASURE_PART_LIST = ['neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'left_feet', 'right_ham',
                     'right_shank', 'right_feet', 'left_arm', 'left_forearm', 'left_hand', 'right_arm',
                     'right_forearm', 'right_hand']
KPS_INDEX_LIST = [[12, 25, 26, 27], [12, 11], [11, 8], [5, 0], [0, 1, 2], [1, 3], [3, 6], [6, 9, 28, 30], [2, 4],
                  [4, 7], [7, 10, 29, 31], [13, 15], [15, 17], [17, 19, 21, 23], [14, 16], [16, 18],
                  [18, 20, 22, 24]]
