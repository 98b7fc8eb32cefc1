def convert_data_format(data_format, ndim):
    return {3: "NWC", 4: "NHWC"}[ndim] if data_format == "channels_last" else {3: "NCW", 4: "NCHW"}[ndim]


def normalize_tuple(value, n, name=None):
    return (int(value),) * n if isinstance(value, int) else tuple(int(v) for v in value)
