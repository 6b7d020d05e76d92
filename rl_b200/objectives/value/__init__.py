from .advantages import GAE, TD0Estimator, TD1Estimator, TDLambdaEstimator
from .functional import (generalized_advantage_estimate, td0_advantage_estimate, td0_return_estimate,
                         td1_advantage_estimate, td1_return_estimate, td_lambda_advantage_estimate,
                         td_lambda_return_estimate, vec_generalized_advantage_estimate, vec_td1_advantage_estimate,
                         vec_td1_return_estimate, vec_td_lambda_advantage_estimate, vec_td_lambda_return_estimate,
                         reward2go, vtrace_advantage_estimate)

__all__ = ["GAE", "TD0Estimator", "TD1Estimator", "TDLambdaEstimator", "generalized_advantage_estimate", "vec_generalized_advantage_estimate", "td0_return_estimate",
           "td0_advantage_estimate", "td1_return_estimate", "vec_td1_return_estimate", "td1_advantage_estimate",
           "vec_td1_advantage_estimate", "td_lambda_return_estimate", "vec_td_lambda_return_estimate",
           "td_lambda_advantage_estimate", "vec_td_lambda_advantage_estimate", "vtrace_advantage_estimate", "reward2go"]
