def multi_gpu_model(model, gpus=None):
    return model
