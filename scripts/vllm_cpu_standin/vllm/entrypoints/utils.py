def cli_env_setup():
    return None
