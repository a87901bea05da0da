from detectron2.modeling.meta_arch.build import META_ARCH_REGISTRY, build_model  # noqa
