class ConfigException(Exception):
    pass


def load_incluster_config():
    raise ConfigException("kubernetes stub: not in a cluster")


def load_kube_config():
    raise ConfigException("kubernetes stub: no kubeconfig")
