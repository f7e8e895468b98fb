// stub: the reference's matrix_market.hh includes this header but uses nothing from it
#pragma once
