// TEST INFRASTRUCTURE: prints the JSON record of one CU exactly as the reference's dataset dumper serialises it
// (Src_HARP/CShow_PredResiReco.h:79-99: CuData builds a Json::Value and streams it with `file << root << endl`).  Built by
// oracle/build_ref.py against the reference's own vendored jsoncpp (Src_HARP/jsoncpp.cpp, compiled where it lies) into
// oracle/_ref/harp_json_ref; tests/golden/make_harp_golden.py turns its output into the committed fixture that pins
// fasthevc_b200/harp_dump.py's writer.  The dumper itself (OpenCV 3.4 C++) cannot be built here.
// usage: harp_json_ref CTUIdx CTUPosX CTUPosY PosX PosY Width Depth AbsPartIdx RastPartIdx RelPosX RelPosY Mode QP POC imgSrc
#include <cstdlib>
#include <iostream>
#include <string>
#include "json.h"

int main(int argc, char** argv)
{
  if (argc != 16) { std::cerr << "15 arguments expected\n"; return 2; }
  Json::Value root;
  root["CTUIdx"] = (Json::UInt)std::strtoul(argv[1], 0, 10);
  root["CTUPos"]["x"] = std::atoi(argv[2]);
  root["CTUPos"]["y"] = std::atoi(argv[3]);
  root["Pos"]["x"] = std::atoi(argv[4]);
  root["Pos"]["y"] = std::atoi(argv[5]);
  root["Width"] = (Json::UInt)std::strtoul(argv[6], 0, 10);
  root["Depth"] = (Json::UInt)std::strtoul(argv[7], 0, 10);
  root["AbsPartIdx"] = (Json::UInt)std::strtoul(argv[8], 0, 10);
  root["RastPartIdx"] = (Json::UInt)std::strtoul(argv[9], 0, 10);
  root["RelPos"]["x"] = std::atoi(argv[10]);
  root["RelPos"]["y"] = std::atoi(argv[11]);
  root["Mode"] = std::string(argv[12]);
  root["QP"] = (Json::UInt)std::strtoul(argv[13], 0, 10);
  root["POC"] = (Json::UInt)std::strtoul(argv[14], 0, 10);
  root["imgSrc"] = std::string(argv[15]) + ".png";
  std::cout << root << std::endl;
  return 0;
}
