def detector_postprocess(results, output_height, output_width):
    raise NotImplementedError("not on the probabilistic-inference path")
