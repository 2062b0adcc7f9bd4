from config.config import ICP_PARAMETERS
