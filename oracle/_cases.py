"""Golden-vector case tables (shared by oracle/make_golden.py and tests/cases.py)."""

CONV_CASES = [
    # name, rank, x_shape, filters, kernel_size, kwargs
    ("c1_same_relu", 1, (2, 19, 12), 5, 3, dict(padding="same", activation="relu")),
    ("c1_valid_s2", 1, (2, 19, 12), 5, 3, dict(padding="valid", strides=2)),
    ("c1_causal_d2", 1, (2, 19, 12), 4, 2, dict(padding="causal", dilation_rate=2, activation="relu")),
    ("c1_same_even_k", 1, (1, 10, 8), 3, 4, dict(padding="same", use_bias=False)),
    ("c1_same_s3", 1, (2, 20, 4), 2, 5, dict(padding="same", strides=3, activation="tanh")),
    ("c1_cf", 1, (2, 12, 17), 4, 3, dict(padding="same", data_format="channels_first", activation="relu")),
    ("c1_decoda_first", 1, (3, 250, 4), 32, 3, dict(padding="same", activation="relu")),
    # shapes the tensor-core kernel takes (in_q % 4 == 0, filters % 16 == 0, stride 1, channels_last)
    ("c1_tc_cfg2_small", 1, (2, 200, 160), 64, 3, dict(padding="same", activation="relu")),
    ("c1_tc_inq12_f16", 1, (3, 150, 48), 16, 3, dict(padding="same", activation="relu")),
    ("c1_tc_valid_k5", 1, (2, 140, 32), 32, 5, dict(padding="valid")),
    ("c1_tc_causal_d2", 1, (2, 131, 32), 16, 2, dict(padding="causal", dilation_rate=2, activation="relu")),
    ("c1_tc_f128", 1, (1, 130, 16), 128, 3, dict(padding="same", activation="relu")),
    ("c1_tc_nobias_sigmoid", 1, (1, 64, 64), 48, 1, dict(padding="same", use_bias=False, activation="sigmoid")),
    ("c2_same_35", 2, (2, 7, 9, 8), 3, (3, 5), dict(padding="same", activation="relu")),
    ("c2_cf_same_35", 2, (2, 8, 7, 9), 3, (3, 5), dict(padding="same", data_format="channels_first", activation="relu")),
    ("c2_valid_s21_d12", 2, (1, 11, 12, 4), 2, (3, 3), dict(padding="valid", strides=(2, 1))),
    ("c2_same_d2", 2, (1, 9, 8, 4), 2, (3, 2), dict(padding="same", dilation_rate=(2, 2), activation="relu")),
    ("c3_same", 3, (1, 4, 5, 6, 4), 2, (2, 3, 3), dict(padding="same", activation="relu")),
    ("c3_cf_valid_s2", 3, (1, 8, 5, 6, 7), 2, (2, 2, 3), dict(padding="valid", strides=(1, 2, 2), data_format="channels_first")),
    # shapes the channels_first tensor-core kernel takes (in_q % 8 == 0, filters % 32 == 0, stride 1, row length % 4 == 0)
    ("c2_tc_cf_same_33", 2, (2, 32, 6, 132), 32, (3, 3), dict(padding="same", data_format="channels_first", activation="relu")),
    ("c2_tc_cf_valid_35_d21", 2, (1, 64, 9, 64), 64, (3, 5), dict(padding="valid", dilation_rate=(2, 1), data_format="channels_first")),
    ("c1_tc_cf_tanh", 1, (2, 32, 200), 32, 3, dict(padding="same", data_format="channels_first", activation="tanh")),
    ("c2_tc_cf_timit_35", 2, (1, 32, 41, 64), 32, (3, 5), dict(padding="same", data_format="channels_first", use_bias=False)),
]

DENSE_CASES = [
    ("d_small", (5, 12), 8, dict(activation="relu")),
    ("d_linear_nobias", (4, 8), 4, dict(use_bias=False)),
    ("d_tc_northstar_small", (300, 160), 256, dict(activation="relu")),
    ("d_tc_f128", (130, 512), 512, dict(activation="relu")),
    ("d_decoda_first", (7, 1000), 512, dict(activation="relu")),
    ("d_tanh", (9, 16), 16, dict(activation="tanh")),
]
