from .xbot_env_math import SyntheticXBotEnv, XBotEnvMath, XBotMathConfig, synthetic_sim_state

__all__ = ["XBotEnvMath", "XBotMathConfig", "SyntheticXBotEnv", "synthetic_sim_state"]
